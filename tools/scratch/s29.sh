cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s29_tests.log
for c in rgbd_1080p mono_4k mono_tum; do
for v in tile3 tile4; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  export ORBX_PYR_TILE_MINRY=8
  t=$(timeout 200 python tools/time_stages.py $c 64 2>&1 | tail -1 | sed 's/.*liborbx_//')
  u=$(timeout 200 python tools/time_total.py $c 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "$t | $u"
done; done 2>&1 | tee gpurun_out/s29.log
