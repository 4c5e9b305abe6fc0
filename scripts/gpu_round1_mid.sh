mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/mid_pytest.log 2>&1; tail -4 gpurun_out/mid_pytest.log
python scripts/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.md; tail -30 gpurun_out/sweep.md
python examples/train_agent.py -e DiscreteSteps-v0 -a ppo --num-envs 512 --rollouts 2 --max-minibatches 40 > gpurun_out/ppo_512.jsonl 2> gpurun_out/ppo_512.err; cat gpurun_out/ppo_512.jsonl; tail -2 gpurun_out/ppo_512.err
python bench.py --steps 5 --warmup 3 > gpurun_out/mid_bench.json 2> gpurun_out/mid_bench.err; python - <<PY
import json
d=json.load(open("gpurun_out/mid_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","rays_per_s")}, d["e2e"]["value"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["roofline_focus"]["launch_ms"], d["roofline_focus"]["frac"], d["cpu_baseline"])
PY
