cd $GRAFT_REPO_ROOT
for v in "$@"; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  r=$(timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1)
  t=$(timeout 200 python tools/time_stages.py rgbd_1080p 64 2>&1 | tail -1 | sed 's/.*rgbd_1080p//')
  echo "$v | $r | $t"
done 2>&1 | tee gpurun_out/s4.log
