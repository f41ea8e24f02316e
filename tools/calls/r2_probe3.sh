#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 tools/probe/tma_copy_probe 512 | tee gpurun_out/tma_copy_probe3.txt
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:probe_kernel --csv --log-file gpurun_out/probe3_ncu.csv tools/probe/tma_copy_probe 512 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/probe3_ncu.csv')) if len(r)>10]
hdr=rows[0]; i_id=hdr.index('ID'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
d={}
for r in rows[1:]:
    d.setdefault(int(r[i_id]),{})[r[i_m]]=r[i_v]
# 23 launches per variant (3 warm + 20): print the 5th of each
ids=sorted(d)
for k in range(0,len(ids),23):
    e=d[ids[min(k+5,len(ids)-1)]]
    print(k//23, e)
P
