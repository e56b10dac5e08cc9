#!/bin/bash
T=${1:-x14}
O=gpurun_out
mkdir -p $O
for c in 1 2 4; do
echo "bg ctas $c"; B200DM_ADAM_BG_CTAS=$c python scripts/opt_cost.py 2>&1 | grep -v "^$" | tee -a $O/${T}_optcost.log
done
