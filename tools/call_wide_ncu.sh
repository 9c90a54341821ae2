set -u
mkdir -p gpurun_out
timeout 120 python tools/prof_flow_large.py 148 > gpurun_out/plain_wide.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flow_field_wide -s 1 -c 1 -o gpurun_out/prof_wide -f \
    python tools/prof_flow_large.py 148 > gpurun_out/ncu_wide.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/plain_wide.log
