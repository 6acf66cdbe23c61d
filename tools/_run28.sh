cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q -k "stft or istft or inverse or encoder or decoder or evaluation or coders" 2>&1 | tail -5
for r in 0 1 2; do echo "B2S_INV_RING=$r"; B2S_INV_RING=$r timeout 300 python tools/kernel_bench.py 2>&1 | grep -i "istft \|stft backward"; done
