cd $GRAFT_REPO_ROOT
for v in 0 12 11 10 8; do
  echo "CTAS_PER_SM=$v: $(ORBX_FAST_CTAS_PER_SM=$v timeout 300 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1)"
done
for c in 3 4; do
  echo "chunks=$c: $(ORBX_DEVICE_CHUNKS=$c timeout 300 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1)"
  echo "chunks=$c cap 11: $(ORBX_FAST_CTAS_PER_SM=11 ORBX_DEVICE_CHUNKS=$c timeout 300 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1)"
done
