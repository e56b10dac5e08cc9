#!/bin/bash
T=${1:-g8}
O=gpurun_out
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > $O/${T}_train.log 2>&1; echo "exit $?" >> $O/${T}_train.log
tail -2 $O/${T}_train.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --workload ddim --steps 3 > $O/${T}_ddim.log 2>&1; echo "exit $?" >> $O/${T}_ddim.log
tail -2 $O/${T}_ddim.log | cut -c1-300
