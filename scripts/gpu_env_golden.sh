mkdir -p gpurun_out
timeout 600 python oracle/gen_golden_env.py gpu > gpurun_out/env_golden.log 2>&1; echo "exit $?" >> gpurun_out/env_golden.log
tail -4 gpurun_out/env_golden.log
cp gpurun_out/golden_gpu/gpu_env_*.npz tests/golden/ 2>/dev/null
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sequences or registry" > gpurun_out/env_pytest.log 2>&1; tail -15 gpurun_out/env_pytest.log
