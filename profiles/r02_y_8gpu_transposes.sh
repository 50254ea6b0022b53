#!/bin/bash
# 8 GPUs, team mode on config 4: variants of the two slab transposes (peer stores from the FFT passes vs a block-copy kernel)
run() { echo "== $*"; env "$@" timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT profiles/team_step_time.py 2>&1 | grep '^{' ; PORT=$((PORT+1)); }
PORT=29600
run X=default
run SWRT_SLAB_B_COPY=1
run SWRT_SLAB_MODE=copy
run SWRT_SLAB_MODE=copy SWRT_SLAB_B_COPY=1
