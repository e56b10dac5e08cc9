#!/bin/bash
T=${1:-x9}
O=gpurun_out
mkdir -p $O
B200DM_GNF_VAR=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -3 $O/${T}_tests.log
for f in 0 1; do
B200DM_GNF_VAR=$f python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_f$f.json > $O/${T}_train_f$f.log 2>&1
B200DM_GNF_VAR=$f python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${T}_ddim_f$f.json > $O/${T}_ddim_f$f.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train*.log") + glob.glob("$O/${T}_ddim*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms"] for n in ("gn_fwd", "rmsnorm_fwd") if n in k})
PY
