#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ramp_launches.csv python tools/ramp_times.py once 512 > gpurun_out/ramp_ncu.txt 2>&1
tail -3 gpurun_out/ramp_ncu.txt
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ramp_launches.csv')) if len(r)>10 and r[0].isdigit()]
# columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section Name, Metric Name, Metric Unit, Metric Value
names=[(r[4].split('(')[0][:70], float(r[-1].replace(',',''))) for r in rows]
print(len(names),'launches')
# split: find PRRT* part = after the last prrtAppendKernel
last=max(i for i,(n,_) in enumerate(names) if 'prrtAppend' in n)
star=names[last+1:]
from collections import OrderedDict
# last full wave of PRRT*: take the final 40 launches
for n,t in star[-45:]: print(f"{t/1000:8.2f} us  {n}")
print('PRRT* launches', len(star), 'sum us', sum(t for _,t in star)/1000)
PY
