cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --config tasnet > gpurun_out/r2f3_bench_tasnet.json 2>gpurun_out/r2f3_bench_tasnet.err; tail -3 gpurun_out/r2f3_bench_tasnet.err; head -c 600 gpurun_out/r2f3_bench_tasnet.json
