mkdir -p gpurun_out/r2g
nvidia-smi -L > gpurun_out/r2g/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 3 > gpurun_out/r2g/bench_n8.json 2> gpurun_out/r2g/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload s2 --steps 3 --warmup 3 > gpurun_out/r2g/bench_s2_n8.json 2> gpurun_out/r2g/bench_s2_n8.err
tail -3 gpurun_out/r2g/bench_n8.err; head -c 300 gpurun_out/r2g/bench_n8.json; echo; tail -3 gpurun_out/r2g/bench_s2_n8.err; head -c 300 gpurun_out/r2g/bench_s2_n8.json
