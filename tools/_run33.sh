cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_pit_pair' -s 1 -c 1 -o gpurun_out/prof_r2_pair python tools/fused_probe.py 2>&1 | tail -2
