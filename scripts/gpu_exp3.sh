#!/bin/bash
# GroupNorm backward variants 1/6/7, forward variant 0/1 (training + DDIM), after the parity tests
T=${1:-x5}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -3 $O/${T}_tests.log
B200DM_GNF_VAR=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "group or gn_" > $O/${T}_tests_f1.log 2>&1; echo "fvar 1 tests exit $?"
for v in 6 7; do
  B200DM_GNB_VAR=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "group or gn_" > $O/${T}_tests_v$v.log 2>&1; echo "var $v tests exit $?"
done
for v in 1 6 7; do
  B200DM_GNB_VAR=$v python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_v$v.json > $O/${T}_train_v$v.log 2>&1
done
B200DM_GNF_VAR=1 python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_f1.json > $O/${T}_train_f1.log 2>&1
for f in 0 1; do
  B200DM_GNF_VAR=$f python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${T}_ddim_f$f.json > $O/${T}_ddim_f$f.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train_*.log") + glob.glob("$O/${T}_ddim_*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms"] for n in ("gn_apply_bwd", "gn_fwd", "rmsnorm_bwd", "linattn_fwd", "rmsnorm_fwd") if n in k})
PY
