#!/bin/bash
# Round 2, GPU call P (2 GPUs): whole GPU suite on a 2-GPU box; the coupled loop (opt-in) against the split loop.
set -x
O=gpurun_out/r02p; mkdir -p $O
CSIM_HALO_TIMEOUT_S=20 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
export CSIM_HALO_TIMEOUT_S=20
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/n1_16384.json 2> $O/n1_16384.err
CSIM_LOOP=coupled timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/n2_16384_coupled.json 2> $O/n2_16384_coupled.err
timeout 300 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > $O/n2_16384_split.json 2> $O/n2_16384_split.err
CSIM_LOOP=coupled timeout 300 $TR bench.py --gpus 2 --tile 8192 --steps 40 --warmup 3 --no-e2e > $O/n2_8192_coupled.json 2> $O/n2_8192_coupled.err
ls -la $O
