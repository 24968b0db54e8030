#!/bin/bash
# Round 2, GPU call M (4 GPUs): the {2,2} decomposition — 4-process parity, the default bench line at N = 4.
set -x
O=gpurun_out/r02m; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tests/mp_parity_worker.py > $O/mp_parity_4.log 2>&1; echo "rc=$?" >> $O/mp_parity_4.log
$TR bench.py --gpus 4 --steps 20 --warmup 3 > $O/n4_16384.json 2> $O/n4_16384.err
$TR bench.py --gpus 4 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n4_8192.json 2> $O/n4_8192.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_16384.json 2> $O/n1_16384.err
ls -la $O
