cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/s38
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/s38/r02_gpu_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/s38/r02_gpu_tests_final.txt
timeout 300 python bench.py --no-cpu-baseline --latency-frames 0 --sustained-s 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench value', round(d['value']), 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'e2e', round(d['e2e']['value']))" | tee -a gpurun_out/s38/r02_gpu_tests_final.txt
