mkdir -p gpurun_out
./tools/rng_microbench > gpurun_out/rng_microbench.txt 2>&1
cat gpurun_out/rng_microbench.txt
timeout 900 python baseline/run_numba_cuda.py --envs 4096 --steps 2 --warmup 1 --out gpurun_out/numba_4096.json > gpurun_out/numba_4096.log 2>&1
tail -2 gpurun_out/numba_4096.log
