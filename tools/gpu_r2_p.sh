#!/bin/bash
# chain kernel: prefetch depth and the launch list of the traceback kernels
for v in libd8.so libd16.so libd32.so; do
  echo "== $v"; for i in 1 2; do GA_B200_LIB=$PWD/globalign_b200/_variants/$v python tools/prof_long.py 100000 | tail -1 | cut -c1-100; done
done
mkdir -p gpurun_out/r2p
GA_B200_LIB=$PWD/globalign_b200/_variants/libd8.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2p/launches_d8.csv python tools/prof_long.py 100000 > /dev/null 2>&1
grep -v "^==" gpurun_out/r2p/launches_d8.csv | python -c "
import csv,sys,collections
r=list(csv.reader(sys.stdin)); h=r[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
c=collections.OrderedDict()
for x in r[1:]:
    if len(x)>vi: c.setdefault(x[ki][:40],[]).append(float(x[vi].replace(',','')))
for k,v in c.items(): print(k, len(v), [round(t/1e6,3) for t in v[-2:]])
"
