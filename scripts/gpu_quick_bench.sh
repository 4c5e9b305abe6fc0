# quick A/B of the tracer on one B200: multi-context parity tests + a short bench (run under gpurun)
TAG=${1:-quick}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "multi_context or specialised or full_benchmark" 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -2 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","rays_per_s")}, d["roofline"]["launch_ms"])
PY
