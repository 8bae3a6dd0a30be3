mkdir -p gpurun_out/r2p
nvidia-smi -L > gpurun_out/r2p/gpus.txt
timeout 400 python bench.py --steps 200 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2p/bench_n1.json 2> gpurun_out/r2p/bench_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29530+N)) bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r2p/bench_n$N.json 2> gpurun_out/r2p/bench_n$N.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 8 --workload s2 --steps 3 --warmup 3 > gpurun_out/r2p/bench_s2_n8.json 2> gpurun_out/r2p/bench_s2_n8.err
for f in gpurun_out/r2p/bench_n1.json gpurun_out/r2p/bench_n2.json gpurun_out/r2p/bench_n4.json gpurun_out/r2p/bench_n8.json gpurun_out/r2p/bench_s2_n8.json; do head -c 220 $f; echo; done
