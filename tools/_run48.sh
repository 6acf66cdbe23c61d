cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dc_gram_frame' -s 1 -c 1 -o gpurun_out/prof_r2_dc python tools/dc_probe.py 2>&1 | tail -2
