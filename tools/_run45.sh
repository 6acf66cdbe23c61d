cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2y_tests.log 2>&1; tail -2 gpurun_out/r2y_tests.log
timeout 300 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; tail -c 200 gpurun_out/r2y_bench.json
for c in pit3 dc tasnet train; do timeout 600 python bench.py --config $c --steps 50 --warmup 5 > gpurun_out/r2y_bench_$c.json 2> gpurun_out/r2y_bench_$c.err; tail -c 150 gpurun_out/r2y_bench_$c.json; echo; done
timeout 600 python bench.py --impl reference-gpu --steps 20 --warmup 3 > gpurun_out/r2y_bench_refgpu.json 2> gpurun_out/r2y_bench_refgpu.err; tail -c 150 gpurun_out/r2y_bench_refgpu.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2y_bench_reference.json 2> gpurun_out/r2y_bench_reference.err; tail -c 150 gpurun_out/r2y_bench_reference.json
timeout 600 python tools/kernel_bench.py --out gpurun_out/r2y_kernels.json > gpurun_out/r2y_kernels.txt 2>&1; cat gpurun_out/r2y_kernels.txt | tail -22
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2y_launches.csv python bench.py --steps 24 --warmup 3 > gpurun_out/r2y_bench_under_ncu.log 2>&1; tail -2 gpurun_out/r2y_launches.csv | cut -c1-200
python __graft_entry__.py 2>&1 | tail -3
