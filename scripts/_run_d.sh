O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "conv" --timeout=300 2>&1 | tail -8
for cfg in "128 32" "256 64"; do set -- $cfg; echo "B=$1 S=$2"; timeout 300 python scripts/conv_microbench.py --what gn --batch $1 --size $2 2>&1 | tail -12; done
echo "== fwd/wgrad micro B=128 S=32"; timeout 300 python scripts/conv_microbench.py --what both --batch 128 --size 32 2>&1 | tail -20
echo "== fwd micro B=256 S=64"; timeout 300 python scripts/conv_microbench.py --what fwd --batch 256 --size 64 2>&1 | tail -20
