cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2v_pit_n2.json 2> gpurun_out/r2v_pit_n2.err; tail -c 900 gpurun_out/r2v_pit_n2.json; grep -c "NCCL INFO" gpurun_out/r2v_pit_n2.err
