#!/bin/bash
# the reference algorithm (oracle) as eager PyTorch on the same GPU: kernel-for-kernel bar (cuDNN / cuBLAS / ATen)
T=${1:-er}
O=gpurun_out
mkdir -p $O
for w in train ddim; do
  python bench.py --impl reference --ref-device cuda --workload $w --steps 5 --warmup 3 > $O/${T}_${w}_fp32.log 2>&1; tail -1 $O/${T}_${w}_fp32.log | cut -c1-260
  python bench.py --impl reference --ref-device cuda --ref-autocast --workload $w --steps 5 --warmup 3 > $O/${T}_${w}_amp.log 2>&1; tail -1 $O/${T}_${w}_amp.log | cut -c1-260
done
