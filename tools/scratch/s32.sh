cd $GRAFT_REPO_ROOT
O=gpurun_out/s32; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -3 | tee $O/tests_default.log
ORBX_LIB=$PWD/tools/ab/liborbx_diag.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -3 | tee $O/tests_diag.log
for lib in orbslam2_with_quadrics_b200/liborbx.so tools/ab/liborbx_diag.so; do
  ORBX_LIB=$PWD/$lib timeout 300 python tools/time_stages.py rgbd_1080p 64 2>&1 | tail -1 | tee -a $O/stages.log
  for H in 1 2 3; do
    ORBX_LIB=$PWD/$lib ORBX_TT_HANDLES=$H timeout 300 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1 | tee -a $O/total.log
  done
done
ORBX_TT_HANDLES=2 ORBX_DEVICE_CHUNKS=1 timeout 300 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1 | tee -a $O/total.log
ORBX_TT_HANDLES=4 ORBX_DEVICE_CHUNKS=1 timeout 300 python tools/time_total.py rgbd_1080p 32 2>&1 | tail -1 | tee -a $O/total.log
ORBX_TT_HANDLES=1 timeout 300 python tools/time_total.py rgbd_1080p 256 2>&1 | tail -1 | tee -a $O/total.log
