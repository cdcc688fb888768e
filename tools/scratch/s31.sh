cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_reloc_init.py -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/s31.log
timeout 300 python tools/time_reloc_init.py rgbd_1080p 16 2>&1 | tail -1 | tee -a gpurun_out/s31.log
timeout 300 python tools/time_reloc_init.py mono_tum 16 2>&1 | tail -1 | tee -a gpurun_out/s31.log
