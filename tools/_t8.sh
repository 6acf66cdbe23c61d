cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 8 "${@:3}" > gpurun_out/$2.json 2>gpurun_out/$2.err; head -c 260 gpurun_out/$2.json; echo; }
run 29511 r2f4_n8_bench --steps 200 --warmup 5
run 29512 r2f4_n8_tasnet --config tasnet --steps 200 --warmup 5
run 29513 r2f4_n8_dc --config dc --steps 200 --warmup 5
grep -h "NCCL INFO.*\(nranks\|NVLS\|comm 0x\)" gpurun_out/r2f4_n8_tasnet.err | head -4 | cut -c1-220
