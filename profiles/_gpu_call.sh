mkdir -p gpurun_out/r2b
echo row rowq64 row40 | python profiles/ab_probe.py 2>&1 | grep -E "^\S+\s+\| S1|rror" > gpurun_out/r2b/ab4.txt
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2b/pytest.log
bash tests/golden/dump_inputs.sh gpurun_out/r2b/inputs > gpurun_out/r2b/dump.log 2>&1
timeout 600 python bench.py --steps 200 --warmup 3 > gpurun_out/r2b/bench.json 2> gpurun_out/r2b/bench.err
cat gpurun_out/r2b/ab4.txt; tail -5 gpurun_out/r2b/pytest.log; tail -3 gpurun_out/r2b/bench.err; head -c 300 gpurun_out/r2b/bench.json
