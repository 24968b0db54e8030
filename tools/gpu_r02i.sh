#!/bin/bash
# Round 2, GPU call I (2 GPUs): the peer path's exchange split into own store kernel | wait for the
# neighbours, with and without a pinned shared-memory carve-out; mp parity on both paths.
set -x
O=gpurun_out/r02i; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "multi_process" > $O/pytest_mp.log 2>&1; echo "pytest rc=$?" >> $O/pytest_mp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
B="bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e"
CSIM_HALO=peer $TR $B > $O/n2_16384_peer.json 2> $O/n2_16384_peer.err
CSIM_HALO=peer CSIM_CARVEOUT=50 $TR $B > $O/n2_16384_peer_carve50.json 2> $O/n2_16384_peer_carve50.err
$TR $B > $O/n2_16384_nccl.json 2> $O/n2_16384_nccl.err
CSIM_HALO=peer $TR $B --tile 8192 --steps 20 > $O/n2_8192_peer.json 2> $O/n2_8192_peer.err
CSIM_HALO=peer CSIM_CARVEOUT=50 $TR $B --tile 8192 --steps 20 > $O/n2_8192_peer_carve50.json 2> $O/n2_8192_peer_carve50.err
ls -la $O
