#!/bin/bash
# Round 2, GPU call C: parity of the staged (TMA → shared memory) sweep, A/B at 16384^2 and 8192^2:
# staged T=4 (default) / staged T=3 / register-only T=3 / round-1 build; ncu capture of the staged sweep.
set -x
O=gpurun_out/r02c; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="python bench.py --tile 16384 --steps 30 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
$B > $O/b16_tbs4.json 2> $O/b16.err
CSIM_TB_MAXT=3 $B > $O/b16_tbs3.json 2>> $O/b16.err
CSIM_TB_KERNEL=reg $B > $O/b16_reg3.json 2>> $O/b16.err
$B > $O/b16_tbs4_2.json 2>> $O/b16.err
CSIM_TB_MAXT=3 $B > $O/b16_tbs3_2.json 2>> $O/b16.err
CSIM_TB_KERNEL=reg $B > $O/b16_reg3_2.json 2>> $O/b16.err
for ch in 128 512; do CSIM_TB_CHUNK=$ch $B > $O/b16_tbs4_ch$ch.json 2>> $O/b16.err; done
B8="python bench.py --tile 8192 --steps 40 --warmup 3 --no-cpu-baseline --no-e2e --no-parity"
$B8 > $O/b8_tbs4.json 2> $O/b8.err
CSIM_TB_KERNEL=reg $B8 > $O/b8_reg3.json 2>> $O/b8.err
# the full default line once (parity + e2e + cpu baseline)
python bench.py > $O/bench_default.json 2> $O/bench_default.err
P="python bench.py --tile 16384 --steps 2 --warmup 1 --inner 12 --no-cpu-baseline --no-e2e --no-parity"
ncu --set full --clock-control none --import-source on -k regex:k_step_tbs -s 4 -c 2 -o $O/tbs4_16384 -f $P > $O/ncu_full.log 2>&1
ncu -i $O/tbs4_16384.ncu-rep --page raw --csv > $O/tbs4_16384_raw.csv 2>/dev/null
ncu -i $O/tbs4_16384.ncu-rep --page source --csv --print-source sass > $O/tbs4_16384_source.csv 2>/dev/null
ls -la $O
