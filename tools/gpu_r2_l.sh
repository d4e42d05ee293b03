#!/bin/bash
# hand-off form x table form: parity (short fuzz) and fill time at 100 kb and 20 kb
for v in libshfl.so libxsmem.so; do
 for pt in 0 1; do
  echo "== $v pair_table=$pt"
  export GA_B200_LIB=$PWD/globalign_b200/_variants/$v GA_LONG_PAIR_TABLE=$pt
  timeout 80 python tests/fuzz/fuzz_long.py 10 47 2>&1 | tail -1 | cut -c1-160
  for L in 100000 100000 20000; do timeout 60 python tools/prof_long.py $L 2>&1 | tail -1 | cut -c1-75; done
 done
done
