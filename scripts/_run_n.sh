O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_ddp_gpu.py -q --timeout=800 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2e_bench_2gpu.log 2> $O/r2e_bench_2gpu.err; echo "exit $?" >> $O/r2e_bench_2gpu.log
tail -3 $O/r2e_bench_2gpu.err
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/r2e_bench_2gpu.log") if l.startswith("{")][0])
s=d.get("secondary") or {}
print("N=2 train", round(d["value"]), "ms", round(d["ms_per_step"],3), "exposed", d.get("exposed_allreduce_ms_per_step"), "ddim", s.get("value"), s.get("ms_per_step"), "weak", (s.get("weak") or {}).get("value"))
P
