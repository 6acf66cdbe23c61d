cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; tail -3 gpurun_out/r2s_tests.log
timeout 300 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; tail -c 600 gpurun_out/r2s_bench.json
