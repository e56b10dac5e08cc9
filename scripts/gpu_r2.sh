#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): GPU test suite, bench (both halves of the metric), reference arm.
#   scripts/gpu_r2.sh TAG [tests|bench|all]
TAG=${1:-r2}
WHAT=${2:-all}
O=gpurun_out
mkdir -p $O
rm -f $O/parity_report.jsonl
if [ "$WHAT" = "tests" ] || [ "$WHAT" = "all" ]; then
  timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 > $O/${TAG}_pytest.log 2>&1; echo "exit $?" >> $O/${TAG}_pytest.log
  tail -25 $O/${TAG}_pytest.log
  cp $O/parity_report.jsonl $O/${TAG}_parity_report.jsonl 2>/dev/null
fi
if [ "$WHAT" = "bench" ] || [ "$WHAT" = "all" ]; then
  timeout 900 python bench.py --steps 20 --warmup 5 --profile-out $O/${TAG}_kernels.json > $O/${TAG}_bench.log 2> $O/${TAG}_bench.err; echo "exit $?" >> $O/${TAG}_bench.log
  tail -c 3000 $O/${TAG}_bench.log; tail -5 $O/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.log 2>&1; echo "exit $?" >> $O/${TAG}_bench_reference.log
  tail -c 1500 $O/${TAG}_bench_reference.log
fi
