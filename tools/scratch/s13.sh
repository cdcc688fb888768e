cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_color.py tests/test_cpp_adapter.py -x -q -m gpu 2>&1 | tail -3
for c in rgbd_1080p mono_tum; do
for v in 1 0; do
  echo "chain=$v $(ORBX_PYR_CHAIN=$v timeout 200 python tools/time_stages.py $c 64 2>&1 | tail -1)"
  echo "chain=$v $(ORBX_PYR_CHAIN=$v timeout 200 python tools/time_total.py $c 64 2>&1 | tail -1)"
  echo "chain=$v $(ORBX_PYR_CHAIN=$v timeout 200 python tools/time_total.py $c 1 2>&1 | tail -1)"
done; done
