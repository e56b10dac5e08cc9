#!/bin/bash
# Round-1 evidence run: GPU tests, both bench workloads, ncu launch list and full-set captures.
set -x
TAG=${1:-r20}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/${TAG}_tests.log
python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/${TAG}_kernels.json > gpurun_out/${TAG}_bench_train.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_bench_train.log
python bench.py --workload ddim --steps 2 --warmup 3 > gpurun_out/${TAG}_bench_ddim.log 2>&1; echo "exit $?" >> gpurun_out/${TAG}_bench_ddim.log
python scripts/profile_step.py --steps 3 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/profile_step.py --steps 3 > gpurun_out/${TAG}_ncu1.log 2>&1
python scripts/profile_step.py --steps 1 > gpurun_out/${TAG}_plain1.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:conv3x3_halo|conv_tc_kernel' -c 30 -o gpurun_out/${TAG}_conv python scripts/profile_step.py --steps 1 > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none -k 'regex:gn_bwd|linattn_bwd|wgrad_tc|rmsnorm_bwd|colsum' -c 40 -o gpurun_out/${TAG}_bwd python scripts/profile_step.py --steps 1 > gpurun_out/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none -k 'regex:gn_stats|gn_apply_fwd|linattn_ctx|linattn_out|rmsnorm_fwd' -c 20 -o gpurun_out/${TAG}_fwd python scripts/profile_step.py --steps 1 > gpurun_out/${TAG}_ncu4.log 2>&1
ls -la gpurun_out
