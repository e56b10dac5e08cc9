#!/bin/bash
# last check of a change: the whole GPU suite, smoke(), the training and DDIM benches
T=${1:-fin}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -4 $O/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
python bench.py --steps 30 --profile-out $O/${T}_kernels.json > $O/${T}_bench_train.log 2>&1; echo "exit $?" >> $O/${T}_bench_train.log
python bench.py --workload ddim --steps 3 --no-cpu-baseline > $O/${T}_bench_ddim.log 2>&1; echo "exit $?" >> $O/${T}_bench_ddim.log
python - <<PY
import json
for f in ("$O/${T}_bench_train.log", "$O/${T}_bench_ddim.log"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); k = d["kernels"]
            print(f.split("/")[-1], round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["launches_per_step"], {n: (k[n]["ms"], k[n]["launches"]) for n in ("conv_tc_fwd", "upsample2x_fwd", "wgrad_tc") if n in k})
PY
