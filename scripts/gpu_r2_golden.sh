# round 2: new goldens from the unmodified reference (numba-CUDA) + the new parity tests (run under gpurun)
mkdir -p gpurun_out/golden_gpu
python oracle/gen_golden_env.py gpu_vis > gpurun_out/r2_golden_vis.log 2>&1; tail -2 gpurun_out/r2_golden_vis.log

cp gpurun_out/golden_gpu/gpu_env_vector_with_renders.npz gpurun_out/golden_gpu/gpu_env_single_with_renders.npz tests/golden/
python -m pytest tests -m gpu -x -q -k "visualizer or binding or generic or sequences or average" > gpurun_out/r2_golden_pytest.log 2>&1; tail -15 gpurun_out/r2_golden_pytest.log
