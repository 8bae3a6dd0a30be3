mkdir -p gpurun_out/r2fb2
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/r2fb2/bench_n1.json 2> gpurun_out/r2fb2/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2fb2/bench_ref.json 2> gpurun_out/r2fb2/bench_ref.err
tail -2 gpurun_out/r2fb2/bench_n1.err; head -c 300 gpurun_out/r2fb2/bench_n1.json; echo; head -c 200 gpurun_out/r2fb2/bench_ref.json; echo
