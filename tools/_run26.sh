cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; tail -3 gpurun_out/r2u_tests.log
timeout 300 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; tail -c 300 gpurun_out/r2u_bench.json
timeout 600 python tools/kernel_bench.py --out gpurun_out/r2u_kernels.json > gpurun_out/r2u_kernels.txt 2>&1; tail -30 gpurun_out/r2u_kernels.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'stft_pit_fused|stft1024_warp' -s 2 -c 2 -o gpurun_out/prof_r2c python tools/fused_probe.py > gpurun_out/r2u_ncu.log 2>&1; tail -1 gpurun_out/r2u_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2u_launches.csv python bench.py --steps 24 --warmup 3 > gpurun_out/r2u_bench_under_ncu.log 2>&1; tail -3 gpurun_out/r2u_launches.csv
