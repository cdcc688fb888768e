cd $GRAFT_REPO_ROOT
timeout 900 python bench.py 2>gpurun_out/s6_bench.err | tee gpurun_out/s6_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'])
print('e2e',{k:v for k,v in d['e2e'].items() if k not in ('call','h2d_ceiling_note')})
print('dropin',d['e2e_dropin']); print('sustained',d['sustained']); print('lat',d['latency_ms']); print('traffic',d['roofline']['traffic'],d['roofline']['traffic_source']); print(d['config']['host_placement'])
"
tail -3 gpurun_out/s6_bench.err
timeout 300 python tools/h2d_ceiling.py | tee gpurun_out/s6_h2d_1gpu.json | cut -c1-600
nvidia-smi topo -m; lscpu | grep -i "numa\|^CPU(s)\|model name"
