cd $GRAFT_REPO_ROOT
timeout 300 python tools/time_stages.py rgbd_1080p 32 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"fast_strips|describe" -s 4 -c 2 -o gpurun_out/r02_fs2 -f python tools/time_stages.py rgbd_1080p 32 > gpurun_out/s16_ncu.log 2>&1
tail -3 gpurun_out/s16_ncu.log
