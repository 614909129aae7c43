#!/bin/bash
# session 2, job M: per-kernel durations of the C-grid subcycle (five-kernel and fused forms), warm caches
mkdir -p gpurun_out
EVP_B200_CGRID_FUSED=0 timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sectors.sum --cache-control none --clock-control none -k regex:k[1-5]_ -s 100 -c 10 --csv --log-file gpurun_out/s2m_cgrid5.csv python scripts/cgrid_time.py 40 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,lts__t_sectors.sum --cache-control none --clock-control none -k regex:k[AB5]_ -s 60 -c 6 --csv --log-file gpurun_out/s2m_cgrid3.csv python scripts/cgrid_time.py 40 > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/s2m_cgrid5.csv","gpurun_out/s2m_cgrid3.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    by={}
    for r in rows:
        by.setdefault((r[0],r[4][:40]),{})[r[12]]=r[14]
    for (i,k),m in by.items():
        print(i,k,{a.split('.')[0][-28:]:b for a,b in m.items()})
PY
