# Profiles the two hot kernels with Nsight Compute on one B200 (run under gpurun).
# usage: bash scripts/gpu_profile.sh <tag> [envs]   -> gpurun_out/<tag>_*
TAG=${1:-prof}
ENVS=${2:-256}
ARGS="--envs $ENVS --steps 2 --warmup 3 --no-cpu-baseline --no-reference-baselines"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:trace_ -s 3 -c 1 -f -o gpurun_out/${TAG}_trace python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:focus -s 3 -c 1 -f -o gpurun_out/${TAG}_focus python bench.py $ARGS > gpurun_out/${TAG}_ncu3.log 2>&1
cut -c1-300 gpurun_out/${TAG}_plain.json; tail -3 gpurun_out/${TAG}_ncu2.log
