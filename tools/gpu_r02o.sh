#!/bin/bash
# Round 2, GPU call O (2 GPUs): the coupled loop with the next sweep queued before the exchange's NCCL enqueue.
set -x
O=gpurun_out/r02o; mkdir -p $O
python -m pytest tests -m gpu -x -q -k "multi_gpu_matches or multi_process_parity_under_torchrun" > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_16384.json 2> $O/n1_16384.err
$TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/n2_16384_coupled.json 2> $O/n2_16384_coupled.err
CSIM_LOOP=split $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/n2_16384_split.json 2> $O/n2_16384_split.err
python bench.py --tile 8192 --steps 40 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_8192.json 2> $O/n1_8192.err
$TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192_coupled.json 2> $O/n2_8192_coupled.err
ls -la $O
