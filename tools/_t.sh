cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f4_bench_reference.json 2>gpurun_out/r2f4_bench_reference.err; head -c 300 gpurun_out/r2f4_bench_reference.json; echo
timeout 300 python bench.py > gpurun_out/r2f4_bench.json 2>gpurun_out/r2f4_bench.err; head -c 300 gpurun_out/r2f4_bench.json; echo
timeout 300 python bench.py --config dc > gpurun_out/r2f4_bench_dc.json 2>/dev/null; head -c 200 gpurun_out/r2f4_bench_dc.json; echo
timeout 300 python bench.py --config pit3 > gpurun_out/r2f4_bench_pit3.json 2>/dev/null; head -c 200 gpurun_out/r2f4_bench_pit3.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f4_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2f4_bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2f4_launches.csv | cut -c1-200
