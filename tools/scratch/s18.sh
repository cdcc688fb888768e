cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/s18_tests.log
for c in rgbd_1080p stereo_kitti mono_tum; do
for v in "$@"; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  t=$(timeout 200 python tools/time_stages.py $c 64 2>&1 | tail -1 | sed 's/.*liborbx_//')
  u=$(timeout 200 python tools/time_total.py $c 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  l=$(timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "$t | $u | B=1: $l"
done; done 2>&1 | tee gpurun_out/s18.log
