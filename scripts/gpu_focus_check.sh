# focus-kernel parity tests + short bench on one B200 (run under gpurun)
TAG=${1:-focus}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "focus or step_focus or gray or full_benchmark" 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["roofline"]["launch_ms"], d["roofline_focus"])
PY
