cd $GRAFT_REPO_ROOT
N=$1
mkdir -p gpurun_out/mg
nvidia-smi topo -m > gpurun_out/mg/topo_${N}gpu.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/mg/r02_bench_1080p_${N}gpu.json 2> gpurun_out/mg/b${N}.err
tail -c 1500 gpurun_out/mg/r02_bench_1080p_${N}gpu.json
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_gpu_multidevice.py tests/test_cpp_adapter.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/mg/r02_gpu_tests_2gpu_box.txt
fi
