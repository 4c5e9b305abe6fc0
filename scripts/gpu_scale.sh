# Multi-GPU bench on N GPUs of one box (run under gpurun --gpus N). usage: bash scripts/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale_${N}_gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/scale_${N}.json 2> gpurun_out/scale_${N}.err; echo "exit $?" >> gpurun_out/scale_${N}.err
tail -3 gpurun_out/scale_${N}.err; cut -c1-400 gpurun_out/scale_${N}.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/scale_${N}_ref.json 2>> gpurun_out/scale_${N}.err
cut -c1-300 gpurun_out/scale_${N}_ref.json
