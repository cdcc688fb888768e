cd $GRAFT_REPO_ROOT
export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_tile2.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/s26_tests.log
for c in rgbd_1080p mono_tum; do
for mr in 32 16 8 4 2; do
  export ORBX_PYR_TILE_MINRY=$mr
  t=$(timeout 200 python tools/time_stages.py $c 64 2>&1 | tail -1 | sed 's/.*liborbx_//')
  u=$(timeout 200 python tools/time_total.py $c 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  l=$(timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "minry=$mr $t | $u | B=1 $l"
done; done 2>&1 | tee gpurun_out/s26.log
