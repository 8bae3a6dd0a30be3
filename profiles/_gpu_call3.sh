mkdir -p gpurun_out/r2n2
nvidia-smi -L > gpurun_out/r2n2/gpus.txt
(timeout 1200 python -m pytest tests -m gpu -x -q -k "comm or shard or multi or group or devices" 2>&1 | tail -15) > gpurun_out/r2n2/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 100 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n2/bench_n2.json 2> gpurun_out/r2n2/bench_n2.err
tail -4 gpurun_out/r2n2/pytest.log; head -c 400 gpurun_out/r2n2/bench_n2.json; tail -2 gpurun_out/r2n2/bench_n2.err
