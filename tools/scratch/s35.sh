cd $GRAFT_REPO_ROOT
O=gpurun_out/s35; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_frustum.py tests/test_gpu_search_projection.py -x -q -m gpu 2>&1 | tail -15 | tee $O/tests.log
