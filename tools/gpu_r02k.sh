#!/bin/bash
# Round 2, GPU call K (1 GPU): the whole GPU suite with the checked constant-divisor division, the default bench
# line, IEEE-division mode at T = 1/2/3 again.
set -x
O=gpurun_out/r02k; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
D="python bench.py --dx 0.3 --dy 0.7 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
$D > $O/div_T1.json 2> $O/div.err
CSIM_TB_DIV_MAXT=2 $D > $O/div_T2.json 2>> $O/div.err
CSIM_TB_DIV_MAXT=3 $D > $O/div_T3.json 2>> $O/div.err
python bench.py > $O/bench_default.json 2> $O/bench_default.err
ls -la $O
