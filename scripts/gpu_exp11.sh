#!/bin/bash
T=${1:-x13}
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q -k "adam or optim or module" > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -3 $O/${T}_tests.log
python scripts/opt_cost.py > $O/${T}_optcost.log 2>&1
cat $O/${T}_optcost.log
python scripts/hbm_microbench.py --out $O/${T}_hbm.json > $O/${T}_hbm.log 2>&1
grep -i adam $O/${T}_hbm.log
