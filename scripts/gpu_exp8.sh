#!/bin/bash
T=${1:-x10}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -5 $O/${T}_tests.log
for f in 1 0; do
B200DM_FUSE_UPSAMPLE=$f python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${T}_ddim_u$f.json > $O/${T}_ddim_u$f.log 2>&1
done
python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_d.json > $O/${T}_train_d.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train*.log") + glob.glob("$O/${T}_ddim*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), d["launches_per_step"], {n: (k[n]["ms"], k[n]["launches"]) for n in ("conv_tc_fwd", "upsample2x_fwd", "gn_fwd") if n in k})
PY
