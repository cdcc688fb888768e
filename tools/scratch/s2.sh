set -x
cd $GRAFT_REPO_ROOT
timeout 200 python tools/time_stages.py rgbd_1080p 16 2>&1 | tail -1 && \
ncu --set full --clock-control none --import-source on -k regex:fast_strips -s 4 -c 2 -o gpurun_out/r02_fast_v5 -f python tools/time_stages.py rgbd_1080p 16 > gpurun_out/ncu_s2.log 2>&1
tail -5 gpurun_out/ncu_s2.log
ls -la gpurun_out/
