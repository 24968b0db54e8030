#!/bin/bash
# Round 2, GPU call D (2 GPUs): the multi-GPU parity tests, weak scaling N = 1 → 2 with the graph-replayed
# block loop, the same loop enqueued eagerly (CSIM_GRAPH=0), 8192^2 tiles, the driver's two-rank file.
set -x
O=gpurun_out/r02d; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/n1_16384.json 2> $O/n1_16384.err
$TR bench.py --gpus 2 --steps 20 --warmup 3 > $O/n2_16384.json 2> $O/n2_16384.err
CSIM_GRAPH=0 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/n2_16384_nograph.json 2> $O/n2_16384_nograph.err
python bench.py --tile 8192 --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_8192.json 2> $O/n1_8192.err
$TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192.json 2> $O/n2_8192.err
CSIM_GRAPH=0 $TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192_nograph.json 2> $O/n2_8192_nograph.err
ls -la $O
