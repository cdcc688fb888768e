cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -4
for B in 64; do
  timeout 200 python tools/time_stages.py rgbd_1080p $B 2>&1 | tail -1
done
timeout 200 python tools/time_stages.py mono_4k 16 2>&1 | tail -1
timeout 200 python tools/time_stages.py stereo_kitti 64 2>&1 | tail -1
timeout 200 python tools/time_stages.py mono_tum 64 2>&1 | tail -1
