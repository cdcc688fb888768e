set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/s1_tests.log
cat gpurun_out/s1_tests.log
for B in 32 64; do
  timeout 200 python tools/time_stages.py rgbd_1080p $B 2>&1 | tail -2
  ORBX_FAST_LEGACY=1 timeout 200 python tools/time_stages.py rgbd_1080p $B 2>&1 | tail -2
done > gpurun_out/s1_stages.log 2>&1
cat gpurun_out/s1_stages.log
