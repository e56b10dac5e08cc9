#!/bin/bash
T=${1:-x15}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -5 $O/${T}_tests.log
python bench.py --steps 30 --no-cpu-baseline --profile-out $O/${T}_train_d.json > $O/${T}_train_d.log 2>&1
python bench.py --workload train64 --steps 20 --no-cpu-baseline --profile-out $O/${T}_train64_d.json > $O/${T}_train64_d.log 2>&1
python scripts/hbm_microbench.py --out $O/${T}_hbm.json > $O/${T}_hbm.log 2>&1
grep -i adam $O/${T}_hbm.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${T}_train*.log")):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), {n: k[n]["ms"] for n in ("attn_fwd", "attn_bwd") if n in k})
PY
