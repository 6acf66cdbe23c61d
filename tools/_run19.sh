cd $GRAFT_REPO_ROOT
timeout 600 python tools/variant_bench.py --fused 0,0 --fwd 0 2>&1 | grep -v "^$"
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:'stft_pit_fused' -s 1 -c 1 python tools/fused_probe.py 2>&1 | grep -E "inst_executed|gpu__time"
