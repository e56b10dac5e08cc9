#!/bin/bash
# GroupNorm-backward variants (B200DM_GNB_VAR) and linear-attention cluster sizes, after the parity tests
T=${1:-x4}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -3 $O/${T}_tests.log
for v in 1 2 3 4 5; do
  B200DM_GNB_VAR=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "group or gn_" > $O/${T}_tests_v$v.log 2>&1; echo "var $v tests exit $?"
done
for v in 0 1 2 3 4 5; do
  B200DM_GNB_VAR=$v python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_v$v.json > $O/${T}_train_v$v.log 2>&1
done
for c in 1 2 4; do
  B200DM_LA_CL=$c python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${T}_ddim_la$c.json > $O/${T}_ddim_la$c.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train_v*.log") + glob.glob("$O/${T}_ddim_la*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms"] for n in ("gn_apply_bwd", "gn_fwd", "linattn_fwd", "rmsnorm_fwd") if n in k})
PY
