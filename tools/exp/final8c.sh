set -x
T=r02f
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
PPP_CHECK_N=400000 timeout 300 $TR --nproc-per-node 8 --master-port 29601 tools/multi_gpu_check.py > gpurun_out/${T}_check8.log 2>&1; tail -3 gpurun_out/${T}_check8.log
PPP_BENCH_VERBOSE=1 timeout 600 $TR --nproc-per-node 8 --master-port 29603 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err; grep "own step" gpurun_out/${T}_bench_8gpu.err | tail -16
CUDA_VISIBLE_DEVICES=0,1,2,3 PPP_BENCH_VERBOSE=1 timeout 400 $TR --nproc-per-node 4 --master-port 29605 bench.py --gpus 4 --steps 20 --warmup 5 --no-cfg3 > gpurun_out/${T}_bench_4gpu.json 2> gpurun_out/${T}_bench_4gpu.err &
CUDA_VISIBLE_DEVICES=4,5 PPP_BENCH_VERBOSE=1 timeout 400 $TR --nproc-per-node 2 --master-port 29606 bench.py --gpus 2 --steps 20 --warmup 5 --no-cfg3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err &
wait
