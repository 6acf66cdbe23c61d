cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_pit_pair|stft1024_warp' -s 2 -c 2 -o gpurun_out/prof_r2_final python tools/fused_probe.py > gpurun_out/r2x_ncu.log 2>&1; tail -1 gpurun_out/r2x_ncu.log
