cd $GRAFT_REPO_ROOT
timeout 300 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; tail -c 1500 gpurun_out/r2w_bench.json
