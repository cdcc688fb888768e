cd $GRAFT_REPO_ROOT
export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_cur.so
for c in 1 2 3 4; do for s in 2 3 4; do
  echo "chunks=$c streams=$s: $(ORBX_KERNEL_STREAMS=$s ORBX_DEVICE_CHUNKS=$c timeout 200 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1 | sed 's/.*chunks=[0-9]*://')"
done; done 2>&1 | tee gpurun_out/s22.log
for c in rgbd_1080p mono_tum; do echo "B=1 $c $(timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1)"; done | tee -a gpurun_out/s22.log
