# ncu --set full of one launch of a trace_ab build (run under gpurun): bash scripts/gpu_ncu_ab.sh <binary> <ctx> <tag>
BIN=$1; CTX=$2; TAG=$3
mkdir -p gpurun_out
tools/bin/$BIN 256 1 $CTX > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:trace_m -s 1 -c 1 -f -o gpurun_out/${TAG} tools/bin/$BIN 256 1 $CTX > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log; cat gpurun_out/${TAG}_plain.log
