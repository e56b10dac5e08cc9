#!/bin/bash
T=${1:-g2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu -k "distributed or two_gpu or ddp or nccl or multi" > $O/${T}_tests.log 2>&1; echo "tests exit $?" >> $O/${T}_tests.log
tail -4 $O/${T}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > $O/${T}_train.log 2>&1; echo "exit $?" >> $O/${T}_train.log
tail -3 $O/${T}_train.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload ddim --steps 3 > $O/${T}_ddim.log 2>&1; echo "exit $?" >> $O/${T}_ddim.log
tail -3 $O/${T}_ddim.log | cut -c1-400
