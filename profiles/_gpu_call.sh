mkdir -p gpurun_out/r2e
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/r2e/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e/launches.csv $B > gpurun_out/r2e/ncu_launch.log 2>&1
K1="python profiles/kernel_target.py s1 5"
$K1 > gpurun_out/r2e/plain_s1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_lcp_kernel -s 3 -c 1 -f -o gpurun_out/r2e/score_s1 $K1 > gpurun_out/r2e/ncu_s1.log 2>&1
K2="python profiles/kernel_target.py s1fit 5"
$K2 > gpurun_out/r2e/plain_s1fit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_lcp_kernel -s 3 -c 1 -f -o gpurun_out/r2e/score_s1fit $K2 > gpurun_out/r2e/ncu_s1fit.log 2>&1
timeout 900 python profiles/pose_sanity.py gpurun_out/r2e/pose_sanity.json > gpurun_out/r2e/pose_sanity.log 2>&1
tail -2 gpurun_out/r2e/plain_s1.log gpurun_out/r2e/plain_s1fit.log gpurun_out/r2e/ncu_s1.log gpurun_out/r2e/ncu_s1fit.log gpurun_out/r2e/ncu_launch.log; tail -3 gpurun_out/r2e/pose_sanity.log | cut -c1-600; ls -la gpurun_out/r2e
