cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/s37
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "level0 or odd_shapes or pinned" 2>&1 | tail -12 | tee gpurun_out/s37/tests.log
