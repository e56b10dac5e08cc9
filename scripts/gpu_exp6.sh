#!/bin/bash
T=${1:-x8}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -5 $O/${T}_tests.log
python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_d.json > $O/${T}_train_d.log 2>&1
python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${T}_ddim_d.json > $O/${T}_ddim_d.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train*.log") + glob.glob("$O/${T}_ddim*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms"] for n in ("attn_fwd", "attn_bwd", "linattn_fwd", "gn_fwd") if n in k})
PY
