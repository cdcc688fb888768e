cd $GRAFT_REPO_ROOT
O=gpurun_out/f2
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee $O/r02_gpu_tests_final.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py 2>$O/bench.err > $O/r02_bench_1080p_b64.json; tail -c 300 $O/r02_bench_1080p_b64.json; echo
timeout 300 python bench.py --handles 1 --no-cpu-baseline --latency-frames 0 --sustained-s 0 2>>$O/bench.err > $O/r02_bench_1080p_b64_one_handle.json
timeout 600 python bench.py --impl reference > $O/r02_bench_reference_arm.json 2>>$O/bench.err
for c in mono_tum stereo_euroc stereo_kitti mono_4k; do
  timeout 400 python bench.py --config $c --no-cpu-baseline > $O/r02_bench_$c.json 2>>$O/bench.err
done
timeout 300 python bench.py --steps 3 --warmup 3 --latency-frames 0 --no-cpu-baseline --e2e-threads 1 --sustained-s 0 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench_b64.csv python bench.py --steps 3 --warmup 3 --latency-frames 0 --no-cpu-baseline --e2e-threads 1 --sustained-s 0 > $O/ncu_launches.log 2>&1
timeout 300 python tools/time_stages.py rgbd_1080p 32 > $O/stages_b32.log 2>&1 && \
ncu --set full --clock-control none --import-source on -s 78 -c 13 -o $O/r02_all_kernels -f python tools/time_stages.py rgbd_1080p 32 > $O/ncu_full.log 2>&1
python tools/traffic_json.py $O/r02_all_kernels.ncu-rep 32 > $O/dominant_kernel_traffic.json 2>$O/traffic.err
python tools/profile_summary.py $O/r02_all_kernels.ncu-rep > $O/r02_all_kernels_ncu_full.txt 2>>$O/traffic.err
python tools/ncu_lines.py $O/r02_all_kernels.ncu-rep "fast_strips_kernel<160, 4>" 60 > $O/r02_fast_strips_by_source_line.txt 2>>$O/traffic.err
python tools/ncu_lines.py $O/r02_all_kernels.ncu-rep describe 45 > $O/r02_describe_by_source_line.txt 2>>$O/traffic.err
python tools/ncu_ops.py $O/r02_all_kernels.ncu-rep "fast_strips_kernel<160, 4>" 30 > $O/r02_fast_strips_opcodes.txt 2>>$O/traffic.err
timeout 300 python tools/e2e_tracking.py rgbd_1080p > $O/r02_e2e_tracking_rgbd_1080p.json 2>>$O/bench.err
timeout 300 python tools/time_frustum.py 16 4000 > $O/r02_is_in_frustum.json 2>>$O/bench.err; cat $O/r02_is_in_frustum.json
ls -la $O
