cd $GRAFT_REPO_ROOT
for v in "$@"; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  t=$(timeout 200 python tools/time_stages.py rgbd_1080p 64 2>&1 | tail -1 | sed 's/.*rgbd_1080p//')
  u=$(timeout 200 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1 | sed 's/.*chunks=2://')
  echo "$v | $t | $u"
done 2>&1 | tee gpurun_out/s17.log
