O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_kernels_gpu.py -q -k "train_py or conv_groupnorm" --timeout=400 2>&1 | tail -4
grep -h train_py $O/parity_report.jsonl | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --profile-out $O/r2d_kernels.json > $O/r2d_bench.log 2> $O/r2d_bench.err; echo "exit $?" >> $O/r2d_bench.log
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2d_bench.log") if l.startswith("{")][0])
s=d.get("secondary") or {}
print("train", round(d["value"]), "ms", round(d["ms_per_step"],3), "ddim", s.get("value"), s.get("ms_per_step"))
k=json.load(open("gpurun_out/r2d_kernels_ddim.json"))
for n,v in k["kernels"].items(): print("  ",n,v)
P
