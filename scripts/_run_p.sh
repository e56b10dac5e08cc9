O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_ddp_gpu.py -q --timeout=500 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary > $O/r2i_bench_2gpu.log 2> $O/r2i_bench_2gpu.err; echo "exit $?" >> $O/r2i_bench_2gpu.log
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2i_bench_2gpu.log") if l.startswith("{")][0])
print("N=2 train", round(d["value"]), "ms", round(d["ms_per_step"],3), "exposed", d.get("exposed_allreduce_ms_per_step"))
P
