cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py tests/test_gpu_stereo.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s24_tests.log
for v in "$@"; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  t=$(timeout 200 python tools/time_stages.py rgbd_1080p 64 2>&1 | tail -1 | sed 's/.*liborbx_//')
  u=$(timeout 200 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "$t | $u"
done 2>&1 | tee gpurun_out/s24.log
