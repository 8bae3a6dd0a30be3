mkdir -p gpurun_out/r2j
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2j/pytest.log
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r2j/bench_n1.json 2> gpurun_out/r2j/bench_n1.err
timeout 900 python profiles/pose_sanity.py gpurun_out/r2j/pose_sanity.json > gpurun_out/r2j/pose_sanity.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j/bench_ref.json 2> gpurun_out/r2j/bench_ref.err
tail -4 gpurun_out/r2j/pytest.log; tail -3 gpurun_out/r2j/bench_n1.err; head -c 250 gpurun_out/r2j/bench_n1.json; echo; tail -3 gpurun_out/r2j/pose_sanity.log | cut -c1-700; head -c 300 gpurun_out/r2j/bench_ref.json
