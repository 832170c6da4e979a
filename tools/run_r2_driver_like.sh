#!/bin/bash
# round 2: both arms launched the way the driver launches them (torchrun, 2 ranks): reference arm (rank 0 alone works), then ours
cd /root/repo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 2>/dev/null | tail -1 | cut -c1-600
