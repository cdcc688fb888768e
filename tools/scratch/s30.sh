cd $GRAFT_REPO_ROOT
for v in base fin base fin; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  timeout 300 python -m pytest tests/test_gpu_search_projection.py -x -q -m gpu -k timing_report 2>&1 | tail -1
  python -c "import json; d=json.load(open('gpurun_out/r01_search_projection.json')); print('$v', d['device_ms_per_batch_incl_staging_h2d'], d['host_wall_ms_per_batch_enqueue_only'], d['e2e_ms_per_batch_with_d2h'])"
done 2>&1 | tee gpurun_out/s30.log
export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_fin.so
timeout 300 python tools/time_reloc_init.py rgbd_1080p 16 2>&1 | tail -1 | tee -a gpurun_out/s30.log
timeout 300 python -m pytest tests/test_gpu_reloc_init.py -x -q -m gpu 2>&1 | tail -2 | tee -a gpurun_out/s30.log
