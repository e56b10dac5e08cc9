#!/bin/bash
# C4 / C5 bench lines + grid-size experiments for the one-pass GroupNorm and the linear attention (DDIM shape)
T=${1:-x1}
mkdir -p gpurun_out
python bench.py --workload ddpm --steps 2 > gpurun_out/${T}_ddpm.log 2>&1
tail -c 1500 gpurun_out/${T}_ddpm.log
python bench.py --workload train64 --steps 20 > gpurun_out/${T}_train64.log 2>&1
tail -c 1500 gpurun_out/${T}_train64.log
for m in 4 8 16; do
  B200DM_GNF_MULT=$m python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out gpurun_out/${T}_ddim_m$m.json > gpurun_out/${T}_ddim_m$m.log 2>&1
  B200DM_GNF_MULT=$m python bench.py --workload train --steps 30 --no-cpu-baseline --profile-out gpurun_out/${T}_train_m$m.json > gpurun_out/${T}_train_m$m.log 2>&1
done
B200DM_LA_CL=2 python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out gpurun_out/${T}_ddim_la2.json > gpurun_out/${T}_ddim_la2.log 2>&1
grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/${T}_*.log | paste - - 
