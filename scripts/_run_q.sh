O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -5
