#!/bin/bash
T=${1:-x7}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -3 $O/${T}_tests.log
python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_d.json > $O/${T}_train_d.log 2>&1
for c in 1 3 4; do
  B200DM_LA_CL=$c python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_la$c.json > $O/${T}_train_la$c.log 2>&1
done
for c in 1 3 4; do
  B200DM_RMSB_MULT=$c python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_rm$c.json > $O/${T}_train_rm$c.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms"] for n in ("linattn_fwd", "linattn_bwd", "rmsnorm_bwd", "final_conv_bwd", "linear_bwd") if n in k})
PY
