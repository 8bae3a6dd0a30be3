mkdir -p gpurun_out/r2w
P="python profiles/pose_latency.py --trace-child"
$P > gpurun_out/r2w/plain_child.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sample_bases|cong_prepare|cong_match" -s 12 -c 4 -f -o gpurun_out/r2w/online $P > gpurun_out/r2w/ncu_online.log 2>&1
tail -3 gpurun_out/r2w/ncu_online.log; ls -la gpurun_out/r2w
