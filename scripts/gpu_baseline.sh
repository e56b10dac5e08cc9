#!/bin/bash
# Evidence run: both bench workloads, HBM / UMMA microbenchmarks, ncu launch list of one eager training step and
# full-set captures of the dominant kernels (raw + source pages exported on the box as CSV; gpurun brings back
# at most 64 MiB, so the .ncu-rep files are dropped when they would not fit).
set -x
TAG=${1:-r40}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench_train.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_train.log
python bench.py --workload ddim --steps 2 --warmup 3 --profile-out $O/${TAG}_kernels_ddim.json > $O/${TAG}_bench_ddim.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_ddim.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_reference.log
python bench.py --workload ddpm --steps 2 > $O/${TAG}_bench_ddpm.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_ddpm.log
python bench.py --workload train64 --steps 20 --profile-out $O/${TAG}_kernels_train64.json > $O/${TAG}_bench_train64.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_train64.log
python scripts/opt_cost.py > $O/${TAG}_opt_cost.log 2>&1
python scripts/hbm_microbench.py --out $O/${TAG}_hbm.json > $O/${TAG}_hbm.log 2>&1
python scripts/umma_rate.py > $O/${TAG}_umma_rate.log 2>&1
python scripts/conv_microbench.py --what both --reps 20 --out $O/${TAG}_conv_micro.json > $O/${TAG}_conv_micro.log 2>&1
python scripts/profile_step.py --steps 3 > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_step.py --steps 3 > $O/${TAG}_ncu1.log 2>&1
python scripts/profile_step.py --steps 1 > $O/${TAG}_plain1.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k 'regex:conv3x3_halo|conv_tc_kernel' -c 12 -o $O/${TAG}_conv python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:wgrad' -c 8 -o $O/${TAG}_wgrad python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none -k 'regex:gn_fwd_cluster|gn_bwd_cluster|linattn|rmsnorm|colsum|adam|im2col' -c 14 -o $O/${TAG}_hbmk python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu4.log 2>&1
# the HBM-bound kernels once more at the DDIM shape (B=256, 64x64): second evaluation, first level
python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_plain_ddim.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:linattn_fwd|gn_fwd|rmsnorm_fwd|im2col7|attn_fwd_tc' -s 63 -c 9 -o $O/${TAG}_hbmk_ddim python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_ncu5.log 2>&1
for r in conv wgrad hbmk hbmk_ddim; do
  ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null
done
ncu -i $O/${TAG}_conv.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/${TAG}_conv_halo_src.csv 2>/dev/null
ncu -i $O/${TAG}_wgrad.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/${TAG}_wgrad_src.csv 2>/dev/null
du -sm $O
for r in hbmk_ddim hbmk conv wgrad; do
  if [ $(du -sm $O | cut -f1) -gt 50 ]; then rm -f $O/${TAG}_$r.ncu-rep; fi
done
ls -la $O
