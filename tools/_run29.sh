cd $GRAFT_REPO_ROOT
B2S_INV_RING=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'istft' -s 1 -c 1 -o gpurun_out/prof_r2_ring2b python tools/istft_probe.py 2>&1 | tail -2
