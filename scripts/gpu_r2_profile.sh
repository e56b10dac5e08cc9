#!/bin/bash
# Round-2 ncu evidence on one B200 (gpurun): launch list of an eager training step, tensor-pipe / DRAM metrics of every
# conv launch of a training step and of a DDIM evaluation, and `--set full` captures of the dominant kernels.
# Each ncu command only runs after the same command exited 0 without ncu.
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__m_xbar2l1tex_read_bytes.sum,launch__registers_per_thread"
python scripts/profile_step.py --steps 3 > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_step.py --steps 3 > $O/${TAG}_ncu1.log 2>&1
python scripts/profile_step.py --steps 2 > $O/${TAG}_plain1.log 2>&1 &&
ncu --metrics $M --clock-control none -k 'regex:conv3x3|conv_tc_kernel|wgrad' -s 230 -c 230 --csv --log-file $O/${TAG}_conv_train_metrics.csv python scripts/profile_step.py --steps 2 > $O/${TAG}_ncu2.log 2>&1
K='regex:attn_fwd|conv3x3|conv_tc|final_conv|gn_fwd|im2col7|linattn|la_kmax|la_ctx|la_mid|la_out|rmsnorm|sgemm|sinusoidal|linear_fwd'
python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_plain_ddim.log 2>&1 &&
NEV=$(grep "eval 1" $O/${TAG}_plain_ddim.log | sed 's/.*, \([0-9]*\) launches/\1/') &&
ncu --metrics $M --clock-control none -k "$K" -s $NEV -c $NEV --csv --log-file $O/${TAG}_eval_ddim_metrics.csv python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_ncu3.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:la_kmax|la_ctx|la_mid|la_out' -c 4 -o $O/${TAG}_lablock python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 1 > $O/${TAG}_ncu7.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:conv3x3_gn' -s 26 -c 3 -o $O/${TAG}_conv_gn_ddim python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_ncu4.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:wgrad3x3_halo|conv3x3_halo' -c 6 -o $O/${TAG}_halo_train python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu5.log 2>&1
ncu --set full --clock-control none -k 'regex:linattn_fwd|gn_bwd_cluster|rmsnorm_bwd' -c 6 -o $O/${TAG}_hbmk python scripts/profile_step.py --steps 1 > $O/${TAG}_ncu6.log 2>&1
for r in conv_gn_ddim halo_train hbmk lablock; do
  ncu -i $O/${TAG}_$r.ncu-rep --page raw --csv > $O/${TAG}_${r}_raw.csv 2>/dev/null
done
ncu -i $O/${TAG}_conv_gn_ddim.ncu-rep --page source --csv --launch-skip 0 --launch-count 1 > $O/${TAG}_conv_gn_src.csv 2>/dev/null
du -sm $O
for r in halo_train hbmk lablock conv_gn_ddim; do
  if [ $(du -sm $O | cut -f1) -gt 55 ]; then rm -f $O/${TAG}_$r.ncu-rep; fi
done
ls -la $O | grep ${TAG}_ | awk '{print $5, $9}'
