mkdir -p gpurun_out/r2d
python profiles/_trace_ycb.py > gpurun_out/r2d/trace_ycb.log 2>&1
timeout 900 python profiles/pose_sanity.py gpurun_out/r2d/pose_sanity.json > gpurun_out/r2d/pose_sanity.log 2>&1
cat gpurun_out/r2d/trace_ycb.log; tail -4 gpurun_out/r2d/pose_sanity.log
