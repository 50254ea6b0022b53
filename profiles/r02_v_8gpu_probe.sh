#!/bin/bash
# 8-GPU probes: concurrent PCIe copies on all GPUs, then the bench at N = 8
for i in 0 1 2 3 4 5 6 7; do CUDA_VISIBLE_DEVICES=$i python profiles/pcie_probe.py h2d,both > gpurun_out/r02v_pcie_$i.log 2>&1 & done
wait
for i in 0 1 2 3 4 5 6 7; do echo "gpu $i: $(tr '\n' '|' < gpurun_out/r02v_pcie_$i.log)"; done
nvidia-smi topo -m | head -14
lscpu | grep -E "NUMA|Socket|Model name"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/r02v_bench_8gpu.json 2> gpurun_out/r02v_bench_8gpu.err
tail -c 3000 gpurun_out/r02v_bench_8gpu.json
