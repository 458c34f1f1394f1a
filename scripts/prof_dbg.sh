cd "${GRAFT_REPO_ROOT:-/root/repo}"
export SF_DEBUG_STAGE=54
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_stage -c 2 -o gpurun_out/dbg54_prof -f python scripts/profile_rollout.py > gpurun_out/dbg54.log 2>&1
tail -2 gpurun_out/dbg54.log
