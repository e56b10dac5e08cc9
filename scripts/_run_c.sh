O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "conv_groupnorm_film_silu_one_launch" --timeout=200 2>&1 | tail -3
for cfg in "128 32" "256 64" "32 64"; do set -- $cfg; echo "B=$1 S=$2"; timeout 300 python scripts/conv_microbench.py --what gn --batch $1 --size $2 2>&1 | tail -12; done
