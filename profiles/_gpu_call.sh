mkdir -p gpurun_out/r2fa
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2fa/pytest.log
P="python profiles/pose_latency.py --trace-child"
$P > gpurun_out/r2fa/plain_child.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2fa/pose_launches.csv $P > gpurun_out/r2fa/ncu_child.log 2>&1
timeout 300 python profiles/pose_latency.py --trace > gpurun_out/r2fa/pose_latency.json 2> gpurun_out/r2fa/pose_latency.err
tail -5 gpurun_out/r2fa/pytest.log; tail -3 gpurun_out/r2fa/ncu_child.log; head -c 400 gpurun_out/r2fa/pose_latency.json
