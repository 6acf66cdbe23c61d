cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 5 > gpurun_out/r2z_pit_n8.json 2> gpurun_out/r2z_pit_n8.err; tail -c 700 gpurun_out/r2z_pit_n8.json; echo; grep -c "NCCL INFO" gpurun_out/r2z_pit_n8.err; grep -m2 "nranks 8\|NVLS" gpurun_out/r2z_pit_n8.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --config tasnet --steps 50 --warmup 5 > gpurun_out/r2z_tasnet_n8.json 2> gpurun_out/r2z_tasnet_n8.err; tail -c 900 gpurun_out/r2z_tasnet_n8.json; echo
true
