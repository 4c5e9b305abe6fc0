# Nsight Compute capture of the focus stencil alone at full batch (run under gpurun).
# usage: bash scripts/gpu_profile_focus.sh <tag> [envs]   -> gpurun_out/<tag>_focus4096.*
TAG=${1:-prof}
ENVS=${2:-4096}
mkdir -p gpurun_out
python scripts/focus_latency.py > gpurun_out/${TAG}_focus_latency.jsonl 2> gpurun_out/${TAG}_focus_latency.err &&
ncu --set full --clock-control none --import-source on -k regex:focus_packed -s 10 -c 1 -f -o gpurun_out/${TAG}_focus${ENVS} python scripts/focus_latency.py --only 300 ${ENVS} > gpurun_out/${TAG}_ncu_focus.log 2>&1
cat gpurun_out/${TAG}_focus_latency.jsonl; tail -3 gpurun_out/${TAG}_ncu_focus.log
