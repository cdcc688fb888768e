cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 1200 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_bounds.py tests/test_cpp_adapter.py tests/test_gpu_search_projection.py -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_2gpu_tests.txt
