#!/bin/bash
# Round 2, GPU call Q (8 GPUs): the coupled loop against the split loop at N = 8.
set -x
O=gpurun_out/r02q; mkdir -p $O
export CSIM_HALO_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e"
CSIM_LOOP=coupled timeout 240 $TR $B > $O/n8_16384_coupled.json 2> $O/n8_16384_coupled.err
timeout 240 $TR $B > $O/n8_16384_split.json 2> $O/n8_16384_split.err
CSIM_LOOP=coupled timeout 240 $TR $B --tile 8192 --steps 20 > $O/n8_8192_coupled.json 2> $O/n8_8192_coupled.err
ls -la $O
