#!/bin/bash
# Round 2, GPU call E (2 GPUs): multi-GPU parity (peer path over CUDA IPC, NCCL path, graph replay), weak
# scaling N = 1 → 2 with the light peer halo against the NCCL halo, both tiles.
set -x
O=gpurun_out/r02e; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "multi or two_ranks" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_16384.json 2> $O/n1_16384.err
$TR bench.py --gpus 2 --steps 20 --warmup 3 > $O/n2_16384_peer.json 2> $O/n2_16384_peer.err
CSIM_HALO=nccl $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/n2_16384_nccl.json 2> $O/n2_16384_nccl.err
python bench.py --tile 8192 --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_8192.json 2> $O/n1_8192.err
$TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192_peer.json 2> $O/n2_8192_peer.err
CSIM_HALO=nccl $TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192_nccl.json 2> $O/n2_8192_nccl.err
CSIM_GRAPH=1 $TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192_peer_graph.json 2> $O/n2_8192_peer_graph.err
ls -la $O
