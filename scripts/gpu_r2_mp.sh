# round 2: GPU tests + bench of the mp tracer at 7 and 8 pixels per thread (run under gpurun)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_mp_pytest.log 2>&1; tail -5 gpurun_out/r2_mp_pytest.log
for K in 7 8; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --contexts $K > gpurun_out/r2_mp_bench_$K.json 2> gpurun_out/r2_mp_bench_$K.err
  python - <<PY
import json
d=json.load(open("gpurun_out/r2_mp_bench_$K.json"))
print("K=$K", round(d["value"]), "env-steps/s", round(d["ms_per_step"],1), "ms/step", round(d["roofline"]["launch_ms"],1), "ms trace", d["env_loop"]["value"], d["env_loop_device"]["value"], d["clocks"])
PY
done
