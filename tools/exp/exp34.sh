TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PPP_CHECK_N=400000 timeout 300 $TR --nproc-per-node 2 --master-port 29601 tools/multi_gpu_check.py > gpurun_out/r02c_check2.log 2>&1; tail -3 gpurun_out/r02c_check2.log
PPP_BENCH_VERBOSE=1 timeout 400 $TR --nproc-per-node 2 --master-port 29606 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg3 > gpurun_out/r02c_bench_2gpu.json 2> gpurun_out/r02c_bench_2gpu.err; tail -c 2500 gpurun_out/r02c_bench_2gpu.json
