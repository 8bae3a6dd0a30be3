mkdir -p gpurun_out/r2fb
timeout 900 python bench.py --steps 200 --warmup 3 > gpurun_out/r2fb/bench_n1.json 2> gpurun_out/r2fb/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2fb/bench_ref.json 2> gpurun_out/r2fb/bench_ref.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2fb/smoke.log 2>&1
tail -2 gpurun_out/r2fb/bench_n1.err; head -c 300 gpurun_out/r2fb/bench_n1.json; echo; head -c 300 gpurun_out/r2fb/bench_ref.json; echo; tail -2 gpurun_out/r2fb/smoke.log
