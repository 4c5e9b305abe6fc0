mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1_smi.txt 2>&1
nproc > gpurun_out/r1_nproc.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r1_smoke.log
timeout 600 python oracle/gen_golden_gpu.py > gpurun_out/r1_golden.log 2>&1; echo "golden exit $?" >> gpurun_out/r1_golden.log
timeout 300 python baseline/run_numba_cuda.py --envs 64 --steps 3 --out gpurun_out/numba_64.json > gpurun_out/r1_numba64.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "bench exit $?" >> gpurun_out/r1_bench.err
tail -5 gpurun_out/r1_pytest.log; cat gpurun_out/r1_smoke.log | tail -3; tail -3 gpurun_out/r1_golden.log; cat gpurun_out/r1_bench.json | cut -c1-600
