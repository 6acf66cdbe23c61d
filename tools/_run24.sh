cd $GRAFT_REPO_ROOT
for e in 0 1 2 3; do echo "WS_EXP=$e"; B2S_FUSED_WS_EXP=$e timeout 300 python tools/variant_bench.py --fused '' --fwd '' --ws 1 2>&1 | grep -v "^$"; done
B2S_FUSED_WS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_pit_ws' -s 2 -c 1 -o gpurun_out/prof_r2_ws python tools/fused_probe.py > gpurun_out/r2t_ncu.log 2>&1; tail -2 gpurun_out/r2t_ncu.log
