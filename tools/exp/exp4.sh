set -x
timeout 600 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_multi.py -x -q 2>&1 | tail -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err
tail -5 gpurun_out/bench2.err
timeout 200 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench2_ref.json 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err
tail -3 gpurun_out/bench1.err
