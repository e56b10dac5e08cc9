#!/bin/bash
# bench + per-launch ncu durations of one eager step (no tests)
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench.log 2>&1; echo "exit $?" >> $O/${TAG}_bench.log
python - <<PY
import json
for l in open("$O/${TAG}_bench.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", round(d["value"]), "img/s  ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "launches", d["launches_per_step"], "mfu", d["roofline"]["step_mfu"])
        for k, v in d["kernels"].items():
            print(f"  {k:20s} {v['ms']:8.3f} ms  x{v['launches']:3d}  {v['share']*100:5.1f}%  {v['tflops']}")
PY
tail -2 $O/${TAG}_bench.log | cut -c1-200
python scripts/profile_step.py --steps 2 > $O/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_step.py --steps 2 > $O/${TAG}_ncu1.log 2>&1
tail -2 $O/${TAG}_ncu1.log
