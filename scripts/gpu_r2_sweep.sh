# round 2: config-3 sweep + small-batch latency on one B200 (run under gpurun)
mkdir -p gpurun_out
python scripts/sweep.py > gpurun_out/r2_sweep.jsonl 2> gpurun_out/r2_sweep.md; tail -26 gpurun_out/r2_sweep.md
python scripts/latency_small_batches.py | tee gpurun_out/r2_latency_small.jsonl
