cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_reloc_init.py tests/test_gpu_search_projection.py tests/test_gpu_match_fuzz.py -x -q -m gpu 2>&1 | tail -25 | tee gpurun_out/s23_tests.log
