cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s5_tests.log
timeout 600 python bench.py 2>gpurun_out/s5_bench.err | tee gpurun_out/s5_bench.json
