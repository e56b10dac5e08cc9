timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "conv" --timeout=300 2>&1 | tail -4
echo "== fwd micro B=128 S=32"; timeout 300 python scripts/conv_microbench.py --what fwd --batch 128 --size 32 2>&1 | grep '"k": 3' | head -8
echo "== fwd micro B=256 S=64"; timeout 300 python scripts/conv_microbench.py --what fwd --batch 256 --size 64 2>&1 | grep '"k": 3' | head -8
python scripts/phase_timing.py --batch 256 --size 64 2>&1 | tail -18 | head -12
