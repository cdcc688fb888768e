cd $GRAFT_REPO_ROOT
O=gpurun_out/s36; mkdir -p $O
nproc
for c in stereo_kitti mono_tum rgbd_1080p; do
for t in 4 6 8; do
  timeout 300 python bench.py --config $c --e2e-threads $t --no-cpu-baseline --latency-frames 0 --sustained-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$c threads=$t value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['e2e']['frac_of_h2d_ceiling'],3), 'ceil', round(d['e2e']['h2d_ceiling_gbs'],1))" | tee -a $O/threads.log
done; done
