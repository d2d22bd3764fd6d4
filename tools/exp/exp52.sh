TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PPP_CHECK_N=200000 timeout 200 $TR --nproc-per-node 2 --master-port 29601 tools/multi_gpu_check.py 2>&1 | tail -2
timeout 300 $TR --nproc-per-node 2 --master-port 29606 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg3 > gpurun_out/r02h_bench_2gpu.json 2> gpurun_out/r02h_bench_2gpu.err; tail -c 400 gpurun_out/r02h_bench_2gpu.json
