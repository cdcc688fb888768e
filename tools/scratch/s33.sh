cd $GRAFT_REPO_ROOT
O=gpurun_out/s33; mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee $O/tests.log
timeout 300 python tools/time_stages.py rgbd_1080p 64 2>&1 | tail -1 | tee -a $O/stages.log
for cfg in "1 2 0" "2 1 0" "3 1 0" "2 1 12" "2 1 10" "2 2 0"; do
  set -- $cfg
  echo "cap=$3 $(ORBX_TT_HANDLES=$1 ORBX_DEVICE_CHUNKS=$2 ORBX_FAST_CTAS_PER_SM=$3 timeout 300 python tools/time_total.py rgbd_1080p 64 2>&1 | tail -1)" | tee -a $O/total.log
done
timeout 600 python bench.py --no-cpu-baseline --latency-frames 200 2>$O/bench.err > $O/bench.json; tail -c 600 $O/bench.json; echo
timeout 300 python bench.py --no-cpu-baseline --latency-frames 0 --handles 1 --sustained-s 0 2>>$O/bench.err > $O/bench_h1.json; head -c 300 $O/bench_h1.json; echo
