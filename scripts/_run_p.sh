O=gpurun_out; mkdir -p $O
for c in 8 4; do
B200DM_COMM_CTAS=$c timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $O/r2g_bench_2gpu_$c.log 2> $O/r2g_bench_2gpu_$c.err; echo "exit $?" >> $O/r2g_bench_2gpu_$c.log
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2g_bench_2gpu_$c.log") if l.startswith("{")][0])
print("ctas $c N=2 train", round(d["value"]), "ms", round(d["ms_per_step"],3), "exposed", d.get("exposed_allreduce_ms_per_step"))
P
tail -2 $O/r2g_bench_2gpu_$c.err | cut -c1-200
done
timeout 600 python -m pytest tests/test_ddp_gpu.py -q --timeout=500 2>&1 | tail -2
