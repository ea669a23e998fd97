#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/r2c51_ncu_msg.csv python tools/msg_bench.py > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2c51_ncu_msg.csv',errors='replace') if l.startswith('"'))]
hdr=rows[0]; ix={n:hdr.index(n) for n in ('ID','Kernel Name','Metric Name','Metric Unit','Metric Value','Grid Size')}
per={}
for r in rows[1:]:
    d=per.setdefault(int(r[ix['ID']]),{'name':r[ix['Kernel Name']].split('(')[0][-28:],'grid':r[ix['Grid Size']]})
    d[r[ix['Metric Name']]]=(r[ix['Metric Value']],r[ix['Metric Unit']])
seen={}
for k in sorted(per):
    d=per[k]
    if 'msg' not in d['name']: continue
    key=(d['name'],d['grid'])
    if key in seen: continue
    seen[key]=1
    print(d['name'],d['grid'],d.get('gpu__time_duration.sum'),d.get('dram__bytes_read.sum'),d.get('dram__bytes_write.sum'),d.get('sm__warps_active.avg.pct_of_peak_sustained_active'),d.get('smsp__inst_executed.sum'))
PY
