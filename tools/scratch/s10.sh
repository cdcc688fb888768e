cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/p
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/p/r02_gpu_tests.txt
timeout 600 python bench.py 2>gpurun_out/p/bench.err > gpurun_out/p/r02_bench_1080p_b64.json; tail -c 400 gpurun_out/p/r02_bench_1080p_b64.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/p/r02_bench_reference_arm.json 2>>gpurun_out/p/bench.err
for c in mono_tum stereo_euroc stereo_kitti mono_4k; do
  timeout 400 python bench.py --config $c --no-cpu-baseline > gpurun_out/p/r02_bench_$c.json 2>>gpurun_out/p/bench.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --latency-frames 0 --no-cpu-baseline --e2e-threads 1 --sustained-s 0 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p/r02_launches_bench_b64.csv python bench.py --steps 3 --warmup 3 --latency-frames 0 --no-cpu-baseline --e2e-threads 1 --sustained-s 0 > gpurun_out/p/ncu_launches.log 2>&1
timeout 300 python tools/time_stages.py rgbd_1080p 32 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -s 78 -c 12 -o gpurun_out/p/r02_all_kernels -f python tools/time_stages.py rgbd_1080p 32 > gpurun_out/p/ncu_full.log 2>&1
ls -la gpurun_out/p
