mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "generic" > gpurun_out/generic_pytest.log 2>&1; tail -30 gpurun_out/generic_pytest.log
