# BASELINE config 5: PPO rollout collection on the vector env over 8 GPUs (run under gpurun --gpus 8)
mkdir -p gpurun_out
for mode in "" "--device-env"; do
  tag=host; [ -n "$mode" ] && tag=device
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
    examples/train_agent.py -e DiscreteSteps-v0 -a ppo --num-envs 4096 --rollouts 2 --max-minibatches 8 $mode \
    > gpurun_out/ppo_8gpu_${tag}.jsonl 2> gpurun_out/ppo_8gpu_${tag}.err
  echo "$tag exit $?"; cut -c1-200 gpurun_out/ppo_8gpu_${tag}.jsonl; tail -2 gpurun_out/ppo_8gpu_${tag}.err
done
