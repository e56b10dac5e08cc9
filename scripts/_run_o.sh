O=gpurun_out; mkdir -p $O
for cfg in "8 fp32" "16 fp32" "8 bf16"; do set -- $cfg
B200DM_COMM_CTAS=$1 B200DM_GRAD_WIRE=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --no-secondary > $O/r2h_bench_8gpu_$1_$2.log 2> $O/r2h_bench_8gpu_$1_$2.err; echo "exit $?" >> $O/r2h_bench_8gpu_$1_$2.log
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2h_bench_8gpu_$1_$2.log") if l.startswith("{")][0])
print("ctas $1 wire $2 N=8 train", round(d["value"]), "ms", round(d["ms_per_step"],3), "exposed", d.get("exposed_allreduce_ms_per_step"))
P
done
B200DM_COMM_CTAS=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --workload train64 > $O/r2h_bench_8gpu_train64.log 2> $O/r2h_bench_8gpu_train64.err
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r2h_bench_8gpu_train64.log") if l.startswith("{")][0])
print("train64 N=8", round(d["value"]), "ms", round(d["ms_per_step"],3), "exposed", d.get("exposed_allreduce_ms_per_step"))
P
