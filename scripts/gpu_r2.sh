#!/bin/bash
# Round-2 runs on one B200 (gpurun).   scripts/gpu_r2.sh TAG [tests|bench|evidence|all]
#   tests    : the GPU test suite (+ parity report)
#   bench    : bench.py as the driver runs it (train + DDIM-50 secondary, CPU and eager-GPU baselines), reference arm
#   evidence : the other BASELINE configs, micro-benchmarks and the per-launch metrics pass of a DDIM evaluation
TAG=${1:-r2}
WHAT=${2:-all}
O=gpurun_out
mkdir -p $O
if [ "$WHAT" = "tests" ] || [ "$WHAT" = "all" ]; then
  rm -f $O/parity_report.jsonl
  timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log
  tail -5 $O/${TAG}_pytest.log
  cp $O/parity_report.jsonl $O/${TAG}_parity_report.jsonl 2>/dev/null
fi
if [ "$WHAT" = "bench" ] || [ "$WHAT" = "all" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "exit $?" >> $O/${TAG}_bench.log
  tail -c 600 $O/${TAG}_bench.log; tail -2 $O/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_reference.log
  tail -c 700 $O/${TAG}_bench_reference.log
fi
if [ "$WHAT" = "evidence" ] || [ "$WHAT" = "all" ]; then
  timeout 900 python bench.py --workload train64 --steps 20 --warmup 5 --no-secondary --profile-out $O/${TAG}_kernels_train64.json > $O/${TAG}_bench_train64.log 2> $O/${TAG}_bench_train64.err; echo "exit $?" >> $O/${TAG}_bench_train64.log
  timeout 900 python bench.py --workload ddpm --steps 2 --no-gpu-baseline > $O/${TAG}_bench_ddpm.log 2> $O/${TAG}_bench_ddpm.err; echo "exit $?" >> $O/${TAG}_bench_ddpm.log
  B200DM_FUSE_GN=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $O/${TAG}_bench_nofuse.log 2> $O/${TAG}_bench_nofuse.err; echo "exit $?" >> $O/${TAG}_bench_nofuse.log
  timeout 300 python scripts/hbm_microbench.py --out $O/${TAG}_hbm.json > $O/${TAG}_hbm.log 2>&1
  timeout 300 python scripts/umma_rate.py > $O/${TAG}_umma_rate.log 2>&1
  for cfg in "128 32" "256 64" "32 64"; do set -- $cfg; echo "== batch $1 size $2"; timeout 300 python scripts/conv_microbench.py --what gn --batch $1 --size $2; done > $O/${TAG}_conv_gn_micro.log 2>&1
  for cfg in "128 32" "256 64"; do set -- $cfg; echo "== batch $1 size $2"; timeout 300 python scripts/conv_microbench.py --what both --batch $1 --size $2; done > $O/${TAG}_conv_micro.log 2>&1
  timeout 120 python scripts/phase_timing.py --batch 256 --size 64 > $O/${TAG}_phase_halo.log 2>&1
  timeout 120 python scripts/phase_timing_gn.py --batch 256 --size 64 > $O/${TAG}_phase_gn.log 2>&1
  timeout 120 python scripts/phase_timing_linattn.py > $O/${TAG}_phase_linattn.log 2>&1
  timeout 200 python scripts/side_cost.py > $O/${TAG}_side_cost.log 2>&1
  timeout 200 python scripts/side_cost.py --size 64 --batch 64 >> $O/${TAG}_side_cost.log 2>&1
  B200DM_FUSE_LINATTN=0 timeout 600 python bench.py --workload ddim --steps 2 --no-cpu-baseline --no-gpu-baseline > $O/${TAG}_bench_ddim_nofuse_linattn.log 2> /dev/null
  M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
  python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_plain_ddim.log 2>&1 &&
  NEV=$(grep "eval 1" $O/${TAG}_plain_ddim.log | sed 's/.*, \([0-9]*\) launches/\1/') &&
  ncu --metrics $M --clock-control none -k 'regex:attn_fwd|conv3x3|conv_tc|final_conv|gn_fwd|im2col7|linattn|la_kmax|la_ctx|la_mid|la_out|rmsnorm|sgemm|sinusoidal|linear_fwd' -s $NEV -c $NEV --csv --log-file $O/${TAG}_eval_ddim_metrics.csv python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_ncu3.log 2>&1
  grep -c b200dm $O/${TAG}_eval_ddim_metrics.csv; cat $O/${TAG}_plain_ddim.log
  ls $O | grep ${TAG}_ | tr '\n' ' '
fi
