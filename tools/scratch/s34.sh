cd $GRAFT_REPO_ROOT
O=gpurun_out/s34; mkdir -p $O
run() { echo "fm=$1 dm=$2 B=$3 H=$4: $(ORBX_FAST_GRID_MULT=$1 ORBX_DESC_GRID_MULT=$2 ORBX_TT_HANDLES=$4 ORBX_DEVICE_CHUNKS=1 timeout 300 python tools/time_total.py rgbd_1080p $3 2>&1 | tail -1)" | tee -a $O/total.log; }
run 1 1 64 2
run 2 1 64 2
run 4 1 64 2
run 8 1 64 2
run 1 4 64 2
run 4 4 64 2
run 1 1 128 2
run 4 4 128 2
