cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/p2
timeout 300 python tools/time_stages.py rgbd_1080p 32 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -s 78 -c 14 -o gpurun_out/p2/r02b_all_kernels -f python tools/time_stages.py rgbd_1080p 32 > gpurun_out/p2/ncu_full.log 2>&1
tail -2 gpurun_out/p2/ncu_full.log
ls -la gpurun_out/p2
