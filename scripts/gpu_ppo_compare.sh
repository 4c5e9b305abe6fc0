# PPO rollout collection, host env vs device env, 512 envs on one B200 (run under gpurun).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_device_env.py -x -q 2>&1 | tail -5
python examples/train_agent.py -e DiscreteSteps-v0 -a ppo --num-envs 512 --rollouts 3 --max-minibatches 8 > gpurun_out/ppo_host_512.jsonl 2> gpurun_out/ppo_host_512.err
python examples/train_agent.py -e DiscreteSteps-v0 -a ppo --num-envs 512 --rollouts 3 --max-minibatches 8 --device-env > gpurun_out/ppo_device_512.jsonl 2> gpurun_out/ppo_device_512.err
cut -c1-220 gpurun_out/ppo_host_512.jsonl gpurun_out/ppo_device_512.jsonl; tail -3 gpurun_out/ppo_device_512.err
