cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"galerkin_ap_kernel|galerkin_rt_kernel" -s 8 -c 2 -o gpurun_out/r2s_galerkin python scripts/prof_nmg.py 100 > gpurun_out/r2s_ncu.log 2>&1; echo rc=$?
tail -3 gpurun_out/r2s_ncu.log
