cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dc_gram_ring|dc_backward_frame' -s 6 -c 2 -f -o gpurun_out/prof_r2_dc_config3 python bench.py --config dc --steps 4 --warmup 3 > gpurun_out/prof_dc_config3.log 2>&1; tail -2 gpurun_out/prof_dc_config3.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'pair_stats_kernel|pair_loss_set' -s 4 -c 2 -f -o gpurun_out/prof_r2_pair_cluster python tools/pair_probe.py > gpurun_out/prof_pair_cluster.log 2>&1; tail -2 gpurun_out/prof_pair_cluster.log | cut -c1-200
