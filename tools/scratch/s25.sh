cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s25_tests.log
export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_tile.so
for c in rgbd_1080p mono_4k; do
for nt in 1 0; do
  if [ $nt = 1 ]; then export ORBX_PYR_NO_TILE=1; else unset ORBX_PYR_NO_TILE; fi
  t=$(timeout 200 python tools/time_stages.py $c 64 2>&1 | tail -1 | sed 's/.*liborbx_//')
  u=$(timeout 200 python tools/time_total.py $c 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "no_tile=$nt $t | $u"
done; done 2>&1 | tee gpurun_out/s25.log
