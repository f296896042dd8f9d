#!/bin/bash
# Collect the round's ncu evidence on one B200 (run under gpurun from the repo root):
#   bash tools/profile_round.sh [tag]      -> gpurun_out/<tag>_*.{log,csv,ncu-rep}
# Every program first runs WITHOUT ncu and must exit 0; numbers printed under ncu are never bench values.
# tools/make_profiles.py turns the captures into the tracked summaries under profiles/.
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_bench_short.log 2> $out/${tag}_bench_short.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > $out/${tag}_ncu_launches.log 2>&1
# the Gram kernel of cfg2 (20000 x 100) and of the cfg3 / cfg5 shape (16000 x 300), one launch = the whole matrix
python tools/i8_once.py 20000 100 > $out/${tag}_i8_once.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gram_i8 -s 1 -c 1 -f -o $out/${tag}_prof_gram_i8 \
    python tools/i8_once.py 20000 100 > $out/${tag}_ncu_gram_i8.log 2>&1
python tools/i8_once.py 16000 300 > $out/${tag}_i8_once_k300.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:gram_i8 -s 1 -c 1 -f -o $out/${tag}_prof_gram_i8_k300 \
    python tools/i8_once.py 16000 300 > $out/${tag}_ncu_gram_i8_k300.log 2>&1
# the per-pair kernel on the cfg4 shape (300 frames of 60 + 300 atoms, -ns 60 -e)
python tools/pair_once.py 300 ns60_e > $out/${tag}_pair_once.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 1 -c 1 -f -o $out/${tag}_prof_pair \
    python tools/pair_once.py 300 ns60_e > $out/${tag}_ncu_pair.log 2>&1
ls -la $out/${tag}_*
