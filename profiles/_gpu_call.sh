mkdir -p gpurun_out/r2f
nvidia-smi -L > gpurun_out/r2f/gpus.txt
(timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_shim_gpu.py -m gpu -q -x 2>&1 | tail -25) > gpurun_out/r2f/pytest2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 3 > gpurun_out/r2f/bench_n2.json 2> gpurun_out/r2f/bench_n2.err
tail -8 gpurun_out/r2f/pytest2.log; tail -5 gpurun_out/r2f/bench_n2.err; head -c 400 gpurun_out/r2f/bench_n2.json
