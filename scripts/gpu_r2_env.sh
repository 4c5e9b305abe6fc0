# round 2: device env generality + focus cap + generic path (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_env_pytest.log 2>&1; tail -6 gpurun_out/r2_env_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_env_bench.json 2> gpurun_out/r2_env_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2_env_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_env_bench.json"))
print({k:d.get(k) for k in ("value","ms_per_step","vs_numba_cuda")}); print("generic", d.get("generic")); print("numba generic", d["numba_cuda_baseline"].get("generic"))
PY
