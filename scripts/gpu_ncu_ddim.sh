#!/bin/bash
# ncu --set full of the HBM-bound kernels at the DDIM shape (B=256, 64x64): linear attention, GroupNorm, RMSNorm, qkv 1x1 conv
T=${1:-n1}
O=gpurun_out
mkdir -p $O
python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${T}_plain.log 2>&1 || { tail -5 $O/${T}_plain.log; exit 1; }
tail -2 $O/${T}_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"linattn_fwd|gn_fwd|rmsnorm_fwd|upsample_fwd|im2col7" -s ${NCU_SKIP:-63} -c ${NCU_COUNT:-9} -o $O/${T}_hbm -f python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${T}_ncu.log 2>&1
tail -3 $O/${T}_ncu.log
ncu -i $O/${T}_hbm.ncu-rep --page raw --csv > $O/${T}_hbm_raw.csv 2>/dev/null
ncu -i $O/${T}_hbm.ncu-rep --page source --csv -k regex:linattn_fwd > $O/${T}_la_src.csv 2>/dev/null
ls -la $O/${T}_*
sz=$(stat -c %s $O/${T}_hbm.ncu-rep); if [ "$sz" -gt 40000000 ]; then rm $O/${T}_hbm.ncu-rep; fi
