cd $GRAFT_REPO_ROOT
for c in rgbd_1080p mono_tum; do
for v in "$@"; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  l=$(timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1 | sed 's/.*chunks=2://')
  l2=$(timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "$v $c B=1: $l | $l2"
done; done 2>&1 | tee gpurun_out/s19.log
