timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "conv_groupnorm_film_silu_one_launch" --timeout=300 2>&1 | tail -12
