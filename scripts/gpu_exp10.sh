#!/bin/bash
T=${1:-x12}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -5 $O/${T}_tests.log
for o in 1 0; do
B200DM_OVERLAP_OPT=$o python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_o$o.json > $O/${T}_train_o$o.log 2>&1
done
B200DM_OVERLAP_OPT=1 python bench.py --workload train64 --steps 20 --no-cpu-baseline > $O/${T}_train64_o1.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["launches_per_step"], d["config"].get("optimizer"))
PY
tail -3 $O/${T}_train_o1.log | cut -c1-300
