timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "conv_groupnorm_film_silu_one_launch" --timeout=200 2>&1 | tail -2
for cfg in "256 64" "128 32"; do set -- $cfg; timeout 300 python scripts/conv_microbench.py --what gn --batch $1 --size $2 2>&1 | grep '"k": 3' | cut -c1-220; done
python scripts/phase_timing_gn.py --batch 256 --size 64 2>&1 | tail -9 | head -5
