#!/bin/bash
# Round 2, GPU call H (8 GPUs): the batched peer push against the NCCL halo path at N = 8, both tiles.
set -x
O=gpurun_out/r02h; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/mp_parity_worker.py > $O/mp_parity_8.log 2>&1; echo "rc=$?" >> $O/mp_parity_8.log
B="bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e"
$TR $B > $O/n8_16384_peer.json 2> $O/n8_16384_peer.err
CSIM_HALO=nccl $TR $B > $O/n8_16384_nccl.json 2> $O/n8_16384_nccl.err
$TR $B --tile 8192 --steps 20 > $O/n8_8192_peer.json 2> $O/n8_8192_peer.err
CSIM_HALO=nccl $TR $B --tile 8192 --steps 20 > $O/n8_8192_nccl.json 2> $O/n8_8192_nccl.err
ls -la $O
