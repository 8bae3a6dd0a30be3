mkdir -p gpurun_out/r2l
for N in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29560+N)) bench.py --gpus $N --steps 200 --warmup 3 > gpurun_out/r2l/bench_n$N.json 2> gpurun_out/r2l/bench_n$N.err
done
for f in gpurun_out/r2l/bench_n4.json gpurun_out/r2l/bench_n8.json; do head -c 220 $f; echo; done
