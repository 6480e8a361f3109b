#!/bin/bash
# Round-2 multi-GPU evidence on one 8-GPU box: strong scaling of the 512-fit C4 sweep through the C ABI's multi-device entry
# point (one host process), the per-process sweep under torchrun, and the weak-scaling bench line at N = 2 / 4 / 8.
O=gpurun_out; T=${1:-r2v}
timeout 600 python scripts/sweep_multi.py 512 L-BFGS 1 2 4 8 > $O/${T}_sweep_multi.jsonl 2> $O/${T}_sweep_multi.err; echo "sweep_multi rc=$?"; cat $O/${T}_sweep_multi.jsonl | cut -c1-200
for N in ${NLIST:-2 4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --no-extras > $O/${T}_bench_${N}gpu.json 2> $O/${T}_bench_${N}gpu.err; echo "bench $N rc=$?"
  tail -1 $O/${T}_bench_${N}gpu.json | cut -c1-160
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 -m gpsurvival_b200.sweep --config C4 --fits 512 --batch 64 > $O/${T}_sweep_torchrun_8gpu.json 2> $O/${T}_sweep_torchrun_8gpu.err; echo "sweep torchrun rc=$?"; tail -1 $O/${T}_sweep_torchrun_8gpu.json | cut -c1-200
