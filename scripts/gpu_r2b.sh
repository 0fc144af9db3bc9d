set -x
cd $GRAFT_REPO_ROOT
timeout 300 python scripts/prof_fused.py 100 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_assemble -s 1 -c 1 -o gpurun_out/prof_fused_v1 -f python scripts/prof_fused.py 100 > gpurun_out/ncu_fused.log 2>&1; echo "ncu rc=$?"; tail -5 gpurun_out/ncu_fused.log
