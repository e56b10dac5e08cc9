#!/bin/bash
# Round-1 evidence run: both bench workloads, ncu launch list and full-set captures (raw pages exported
# on the box as CSV; gpurun only brings back <= 64 MiB).
set -x
TAG=${1:-r20}
O=gpurun_out
mkdir -p $O
python bench.py --steps 20 --warmup 5 --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench_train.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_train.log
python bench.py --workload ddim --steps 2 --warmup 3 > $O/${TAG}_bench_ddim.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_ddim.log
python scripts/profile_step.py --steps 3 > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_step.py --steps 3 > $O/${TAG}_ncu1.log 2>&1
python scripts/profile_step.py --steps 1 > $O/${TAG}_plain1.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:conv3x3_halo|conv_tc_kernel' -c 16 -o $O/${TAG}_conv python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none -k 'regex:wgrad_tc' -c 8 -o $O/${TAG}_wgrad python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu3.log 2>&1
for r in conv wgrad; do
  ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null
done
du -sm $O
if [ $(du -sm $O | cut -f1) -gt 55 ]; then rm -f $O/${TAG}_conv.ncu-rep; fi
ls -la $O
