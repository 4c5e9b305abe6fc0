# round 2, first call: pipe microbench, baseline tracer A/B harness, GPU tests (run under gpurun)
mkdir -p gpurun_out
./tools/bin/pipe_microbench > gpurun_out/r2_pipes.txt 2>&1; cat gpurun_out/r2_pipes.txt
for B in tools/bin/trace_ab_*; do echo "== $B"; $B 1024 5 7 | tee -a gpurun_out/r2_trace_ab.jsonl; done
python -m pytest tests -m gpu -x -q > gpurun_out/r2_first_pytest.log 2>&1; tail -3 gpurun_out/r2_first_pytest.log
