mkdir -p gpurun_out/r2/mg1
timeout 500 python -m pytest tests/test_gpu_long_keys.py -q --timeout 120 > gpurun_out/r2/mg1/longkeys.log 2>&1
timeout 500 python -m pytest tests/test_multi_gpu.py -q --timeout 200 > gpurun_out/r2/mg1/multi.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -q --timeout 120 -k "cast or nan or NaN" > gpurun_out/r2/mg1/parity_cast.log 2>&1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/mg1/bench2.json 2> gpurun_out/r2/mg1/bench2.err
tail -5 gpurun_out/r2/mg1/*.log; tail -c 1500 gpurun_out/r2/mg1/bench2.json; tail -5 gpurun_out/r2/mg1/bench2.err
