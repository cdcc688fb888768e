cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stereo.py tests/test_gpu_fuzz.py tests/test_gpu_bounds.py -x -q -m gpu 2>&1 | tail -2
ORBX_FAST_LEGACY=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "configs_stage or adversarial" 2>&1 | tail -1
for c in rgbd_1080p stereo_kitti mono_tum stereo_euroc; do
  timeout 400 python bench.py --config $c --no-cpu-baseline --sustained-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$c', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'lat', round(d['latency_ms']['p50'],4), round(d['latency_ms']['p99'],4), 'dropin', round(d['e2e_dropin']['p50_ms'],3), 'launches', d['gpu_launches'])"
done
