#!/bin/bash
# Round 2, GPU call F (1 GPU): the whole GPU suite, the default bench line and the reference arm, the cost of
# the interior/frame split by itself, IEEE-division mode at T = 1/2/3, the drop-in loop modes, launch list.
set -x
O=gpurun_out/r02f; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
python bench.py > $O/bench_default.json 2> $O/bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
B="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
$B > $O/b16.json 2> $O/b16.err
CSIM_DEBUG_SPLIT=1 $B > $O/b16_split.json 2>> $O/b16.err
D="python bench.py --dx 0.3 --dy 0.7 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e"
$D > $O/div_T1.json 2> $O/div.err
CSIM_TB_DIV_MAXT=2 $D > $O/div_T2.json 2>> $O/div.err
CSIM_TB_DIV_MAXT=3 $D > $O/div_T3.json 2>> $O/div.err
climate-sim-mpi-cpp_b200/host/build/bench_loop_modes 4096 20 > $O/loop_modes.txt 2>&1
climate-sim-mpi-cpp_b200/host/build/bench_loop_modes 16384 6 >> $O/loop_modes.txt 2>&1
P="python bench.py --steps 2 --warmup 1 --inner 12 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_16384.csv $P > $O/ncu_list.log 2>&1
ls -la $O
