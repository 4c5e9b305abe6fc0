mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/mc_pytest.log 2>&1; tail -4 gpurun_out/mc_pytest.log
for K in 0 2 4 8; do
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --contexts $K > gpurun_out/mc_bench_$K.json 2> gpurun_out/mc_bench_$K.err
  python - <<PY
import json
d=json.load(open("gpurun_out/mc_bench_$K.json"))
print("K=$K", round(d["value"]), "env-steps/s", round(d["roofline"]["launch_ms"],1), "ms trace", round(d["rays_per_s"]/1e9,1), "Grays/s")
PY
done
