#!/bin/bash
T=${1:-fin2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -3 $O/${T}_tests.log
B200DM_GNF_THREADS=128 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -k "group or gn_ or unet or config1" > $O/${T}_tests128.log 2>&1; echo "tests128 exit $?"; tail -2 $O/${T}_tests128.log
for t in 256 128; do
B200DM_GNF_THREADS=$t python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_t$t.json > $O/${T}_train_t$t.log 2>&1
B200DM_GNF_THREADS=$t python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${T}_ddim_t$t.json > $O/${T}_ddim_t$t.log 2>&1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train_t*.log") + glob.glob("$O/${T}_ddim_t*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), {n: k[n]["ms"] for n in ("gn_fwd", "gn_apply_bwd") if n in k})
PY
