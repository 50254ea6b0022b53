#!/bin/bash
# 8 GPUs, team mode on config 4 after the slab x-pass address arithmetic change: one run of profiles/team_step_time.py
timeout 80 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 profiles/team_step_time.py 2>&1 | grep '^{'
