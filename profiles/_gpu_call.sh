mkdir -p gpurun_out/r2fb3
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/r2fb3/bench_n1.json 2> gpurun_out/r2fb3/bench_n1.err
tail -2 gpurun_out/r2fb3/bench_n1.err; head -c 300 gpurun_out/r2fb3/bench_n1.json; echo
