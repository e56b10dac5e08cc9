#!/bin/bash
# quick GPU check: parity tests, training bench and DDIM bench with the per-kernel tables
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -x -q > $O/${TAG}_tests.log 2>&1; echo "tests exit $?" >> $O/${TAG}_tests.log
tail -5 $O/${TAG}_tests.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench.log 2>&1; echo "exit $?" >> $O/${TAG}_bench.log
timeout 300 python bench.py --workload ddim --steps 3 --no-cpu-baseline --profile-out $O/${TAG}_kernels_ddim.json > $O/${TAG}_bench_ddim.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_ddim.log
python - <<PY
import json
for f in ("$O/${TAG}_bench.log", "$O/${TAG}_bench_ddim.log"):
  for l in open(f):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", round(d["value"], 1), "img/s  ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d["launches_per_step"], "mfu", d["roofline"]["step_mfu"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        for k, v in d["kernels"].items():
            print(f"  {k:20s} {v['ms']:8.3f} ms  x{v['launches']:3d}  {v['share']*100:5.1f}%  {v['tflops']}")
PY
tail -3 $O/${TAG}_bench_ddim.log | cut -c1-300
