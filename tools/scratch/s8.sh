cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02_topo_8gpu.txt 2>&1; lscpu | grep -i "numa\|^CPU(s)\|model name\|socket" >> gpurun_out/r02_topo_8gpu.txt
cat gpurun_out/r02_topo_8gpu.txt | head -30
for N in 8 4 2; do
  timeout 200 $TR --nproc-per-node $N --master-port $((29600+N)) tools/h2d_ceiling.py --gpus $N 2>gpurun_out/h2d_$N.err | tee gpurun_out/r02_h2d_ceiling_${N}gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('h2d', d['n_gpus'], {k:(v['aggregate_gbs'], v['per_rank_gbs']) for k,v in d['result'].items()}, d['cpus_allowed'], d['numa_cpulists'])"
done
timeout 600 $TR --nproc-per-node 8 --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/b8_auto.err | tee gpurun_out/r02_bench_1080p_8gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('N8 auto value',d['value'],'e2e',e['value'],'threads',e['host_threads'],'ceil',e['h2d_ceiling_gbs'],'frac',e['frac_of_h2d_ceiling']); print([ (x['config']['workload'][:12], x['value'], x['e2e']['value'], x['roofline']['frac']) for x in d.get('extra_configs',[])]); print(d['config']['host_placement'])"
for T in 1 4; do
timeout 400 $TR --nproc-per-node 8 --master-port $((29710+T)) bench.py --gpus 8 --steps 20 --warmup 5 --e2e-threads $T --no-extra-configs --sustained-s 0 --latency-frames 0 2>gpurun_out/b8_t$T.err | tee gpurun_out/r02_bench_1080p_8gpu_t$T.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('N8 value',d['value'],'e2e',e['value'],'threads',e['host_threads'],'ceil',e['h2d_ceiling_gbs'],'frac',e['frac_of_h2d_ceiling'])"
done
timeout 400 $TR --nproc-per-node 8 --master-port 29720 bench.py --gpus 8 --steps 20 --warmup 5 --no-numa-bind --no-extra-configs --sustained-s 0 --latency-frames 0 2>gpurun_out/b8_nobind.err | tee gpurun_out/r02_bench_1080p_8gpu_nobind.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('N8 nobind value',d['value'],'e2e',e['value'],'threads',e['host_threads'],'ceil',e['h2d_ceiling_gbs'],'frac',e['frac_of_h2d_ceiling'])"
timeout 400 $TR --nproc-per-node 4 --master-port 29730 bench.py --gpus 4 --steps 20 --warmup 5 --sustained-s 0 --latency-frames 0 2>gpurun_out/b4.err | tee gpurun_out/r02_bench_1080p_4gpu.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('N4 value',d['value'],'e2e',e['value'],'threads',e['host_threads'],'ceil',e['h2d_ceiling_gbs'],'frac',e['frac_of_h2d_ceiling']); print([ (x['config']['workload'][:12], x['value'], x['e2e']['value'], x['roofline']['frac']) for x in d.get('extra_configs',[])])"
tail -3 gpurun_out/b8_auto.err
