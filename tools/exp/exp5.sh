set -x
timeout 600 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_multi.py -x -q 2>&1 | tail -30
PPP_BENCH_VERBOSE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 20 --warmup 3 --no-cfg3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
grep -v "^\*\|OMP_NUM" gpurun_out/bench2.err | tail -8
