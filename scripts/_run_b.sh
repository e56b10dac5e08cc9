O=gpurun_out; mkdir -p $O; rm -f $O/parity_report.jsonl
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "conv_groupnorm_film_silu_one_launch" --timeout=300 > $O/r2b_gnk.log 2>&1; echo "exit $?" >> $O/r2b_gnk.log
tail -30 $O/r2b_gnk.log
if grep -q "exit 0" $O/r2b_gnk.log; then
  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/r2b_pytest.log 2>&1; echo "exit $?" >> $O/r2b_pytest.log
  tail -15 $O/r2b_pytest.log
  cp $O/parity_report.jsonl $O/r2b_parity_report.jsonl
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --profile-out $O/r2b_kernels.json > $O/r2b_bench.log 2> $O/r2b_bench.err; echo "exit $?" >> $O/r2b_bench.log
  B200DM_FUSE_GN=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $O/r2b_bench_nofuse.log 2> $O/r2b_bench_nofuse.err; echo "exit $?" >> $O/r2b_bench_nofuse.log
  python - <<'P'
import json
for f in ("gpurun_out/r2b_bench.log","gpurun_out/r2b_bench_nofuse.log"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][0])
        print(f, "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), "ddim", round(d["secondary"]["value"],1), d["secondary"]["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
P
fi
