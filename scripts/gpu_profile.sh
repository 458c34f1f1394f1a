#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_rollout.py > gpurun_out/${TAG}_launches.log 2>&1
tail -3 gpurun_out/${TAG}_launches.log
timeout 1200 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_stage \
    -o gpurun_out/${TAG}_prof -f python scripts/profile_rollout.py > gpurun_out/${TAG}_prof.log 2>&1
tail -3 gpurun_out/${TAG}_prof.log
ls -la gpurun_out/
python scripts/ncu_summary.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_ncu_summary.txt > /dev/null 2>&1; tail -14 gpurun_out/${TAG}_ncu_summary.txt
python scripts/traffic_from_ncu.py gpurun_out/${TAG}_prof.ncu-rep gpurun_out/${TAG}_roofline_traffic.json 200 8 | tail -12
