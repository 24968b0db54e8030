#!/bin/bash
# Round 2, GPU call B: parity again, memcheck on the small parity cases, A/B against the round-1 build,
# chunk heights, launch list + ncu full capture at 16384^2.
set -x
O=gpurun_out/r02b; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
B="python bench.py --tile 16384 --steps 30 --warmup 3 --no-cpu-baseline --no-e2e"
CSIM_LIB_PATH=$PWD/ab/libcsim_r01.so $B > $O/b16_r01.json 2> $O/b16_r01.err
$B > $O/b16_new.json 2> $O/b16_new.err
CSIM_LIB_PATH=$PWD/ab/libcsim_r01.so $B > $O/b16_r01_2.json 2>> $O/b16_r01.err
$B > $O/b16_new_2.json 2>> $O/b16_new.err
for ch in 96 192 384 512; do CSIM_TB_CHUNK=$ch $B > $O/b16_new_ch$ch.json 2>> $O/b16_new.err; done
B8="python bench.py --tile 8192 --steps 40 --warmup 3 --no-cpu-baseline --no-e2e"
CSIM_LIB_PATH=$PWD/ab/libcsim_r01.so $B8 > $O/b8_r01.json 2> $O/b8.err
$B8 > $O/b8_new.json 2>> $O/b8.err
CSIM_TB_CHUNK=96 $B8 > $O/b8_new_ch96.json 2>> $O/b8.err
P="python bench.py --tile 16384 --steps 2 --warmup 1 --inner 12 --no-cpu-baseline --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:k_step_tb -s 4 -c 2 -o $O/tb3_16384 -f $P > $O/ncu_full.log 2>&1
ncu -i $O/tb3_16384.ncu-rep --page raw --csv > $O/tb3_16384_raw.csv 2>/dev/null
ncu -i $O/tb3_16384.ncu-rep --page source --csv --print-source sass > $O/tb3_16384_source.csv 2>/dev/null
ls -la $O
