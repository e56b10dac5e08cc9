timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "conv" --timeout=300 2>&1 | tail -2
for cfg in "256 64" "128 32"; do set -- $cfg; timeout 300 python scripts/conv_microbench.py --what gn --batch $1 --size $2 2>&1 | grep '"k": 3' | cut -c1-220 | head -4; done
echo "== fwd/wgrad B=128 S=32"; timeout 300 python scripts/conv_microbench.py --what both --batch 128 --size 32 2>&1 | cut -c1-200 | head -12
echo "== fwd B=256 S=64"; timeout 300 python scripts/conv_microbench.py --what fwd --batch 256 --size 64 2>&1 | cut -c1-200| head -8
