#!/bin/bash
# Round 2, GPU call G (8 GPUs): weak scaling at N = 8 with the default bench line (16384^2 per GPU: parity
# windows on every rank, halo timeline, e2e, shared snapshot file) and at 8192^2 per GPU; 8-process parity.
set -x
O=gpurun_out/r02g; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 8 --steps 20 --warmup 3 > $O/n8_16384.json 2> $O/n8_16384.err
$TR bench.py --gpus 8 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n8_8192.json 2> $O/n8_8192.err
timeout 300 $TR tests/mp_parity_worker.py > $O/mp_parity_8.log 2>&1; echo "rc=$?" >> $O/mp_parity_8.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_16384.json 2> $O/n1_16384.err
ls -la $O
