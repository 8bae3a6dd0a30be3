mkdir -p gpurun_out/r2fin2
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30) > gpurun_out/r2fin2/pytest.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/r2fin2/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2fin2/launches.csv $B > gpurun_out/r2fin2/ncu_launch.log 2>&1
K1="python profiles/kernel_target.py s1 5"
$K1 > gpurun_out/r2fin2/plain_s1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_lcp_kernel -s 3 -c 1 -f -o gpurun_out/r2fin2/score_s1 $K1 > gpurun_out/r2fin2/ncu_s1.log 2>&1
K2="python profiles/kernel_target.py s1fit 5"
$K2 > gpurun_out/r2fin2/plain_s1fit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_lcp_kernel -s 3 -c 1 -f -o gpurun_out/r2fin2/score_s1fit $K2 > gpurun_out/r2fin2/ncu_s1fit.log 2>&1
cp gpurun_out/r2fin/pose_launches.csv gpurun_out/r2fin2/ 2>/dev/null
tail -4 gpurun_out/r2fin2/pytest.log; cat gpurun_out/r2fin2/plain_s1.log gpurun_out/r2fin2/plain_s1fit.log; ls gpurun_out/r2fin2
