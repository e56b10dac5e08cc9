O=gpurun_out; TAG=r2z
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "exit $?" >> $O/${TAG}_bench.log
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2z_bench.log") if l.startswith("{")][0])
s=d["secondary"]
print("TRAIN", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],4), "DDIM", round(s["value"],1), round(s["ms_per_step"],1), s["clocks"])
P
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread"
python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_plain_ddim.log 2>&1 &&
NEV=$(grep "eval 1" $O/${TAG}_plain_ddim.log | sed 's/.*, \([0-9]*\) launches/\1/') &&
ncu --metrics $M --clock-control none -k 'regex:attn_fwd|conv3x3|conv_tc|final_conv|gn_fwd|im2col7|linattn|rmsnorm|sgemm|sinusoidal' -s $NEV -c $NEV --csv --log-file $O/${TAG}_eval_ddim_metrics.csv python scripts/profile_step.py --workload eval --batch 256 --size 64 --steps 2 > $O/${TAG}_ncu3.log 2>&1
grep -c "gpu__time_duration" $O/${TAG}_eval_ddim_metrics.csv
