#!/bin/bash
# Round 2, GPU call L (1 GPU): IEEE-division mode with the branch-free checked division.
set -x
O=gpurun_out/r02l; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
D="python bench.py --dx 0.3 --dy 0.7 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$D > $O/div_T1.json 2> $O/div.err
CSIM_TB_DIV_MAXT=2 $D > $O/div_T2.json 2>> $O/div.err
python bench.py --dx 0.3 --dy 0.7 --tile 8192 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > $O/div_T1_8192.json 2>> $O/div.err
ls -la $O
