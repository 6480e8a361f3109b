#!/bin/bash
# Weak-scaling bench lines on one 8-GPU box (one process per GPU, no collective on the data path).
O=gpurun_out; T=${1:-r3p}
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N bench.py --gpus $N --no-extras > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err; echo "bench $N rc=$?"
  tail -1 $O/${T}_bench_${N}gpu.json | cut -c1-170
done
