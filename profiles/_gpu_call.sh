mkdir -p gpurun_out/r2s
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2s/pytest.log
P="python profiles/pose_latency.py --trace-child"
$P > gpurun_out/r2s/plain_child.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2s/pose_launches.csv $P > gpurun_out/r2s/ncu_child.log 2>&1
tail -5 gpurun_out/r2s/pytest.log; tail -3 gpurun_out/r2s/ncu_child.log; wc -l gpurun_out/r2s/pose_launches.csv
