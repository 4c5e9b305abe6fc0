# A/B of tracer builds under tools/bin (run under gpurun): bash scripts/gpu_ab.sh "<binary args>..." 
mkdir -p gpurun_out
while read -r line; do
  [ -z "$line" ] && continue
  echo "== $line"; timeout 300 tools/bin/$line | tee -a gpurun_out/r2_trace_ab.jsonl
done < scripts/gpu_ab.list
