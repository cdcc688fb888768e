cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py tests/test_gpu_stereo.py -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/s27_tests.log
for c in rgbd_1080p mono_tum; do
for mr in 32 4; do
  export ORBX_PYR_TILE_MINRY=$mr
  t=$(timeout 200 python tools/time_stages.py $c 64 2>&1 | tail -1 | sed 's/.*rgbd_1080p//;s/.*mono_tum//')
  u=$(timeout 200 python tools/time_total.py $c 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  l=$(timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "$c minry=$mr $t | $u | B=1 $l"
done; done 2>&1 | tee gpurun_out/s27.log
