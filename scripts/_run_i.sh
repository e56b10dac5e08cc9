O=gpurun_out; mkdir -p $O; rm -f $O/parity_report.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > $O/r2c_pytest.log 2>&1; echo "exit $?" >> $O/r2c_pytest.log
tail -6 $O/r2c_pytest.log
cp $O/parity_report.jsonl $O/r2c_parity_report.jsonl
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --profile-out $O/r2c_kernels.json > $O/r2c_bench.log 2> $O/r2c_bench.err; echo "exit $?" >> $O/r2c_bench.log
timeout 600 python bench.py --workload train64 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline > $O/r2c_bench_train64.log 2> $O/r2c_bench64.err; echo "exit $?" >> $O/r2c_bench_train64.log
python - <<'P'
import json
for f in ("gpurun_out/r2c_bench.log","gpurun_out/r2c_bench_train64.log"):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][0])
        s=d.get("secondary") or {}
        print(f, "train", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), "launches", d["launches_per_step"], "ddim", s.get("value"), s.get("ms_per_step"), "roof", d["roofline"]["kernel"], d["roofline"]["frac"], "mfu", d["roofline"]["step_mfu"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 $O/r2c_bench.err
