#!/bin/bash
# ncu --set full captures of K9 (linear_tc_kernel) and K10 (shortcut_tc_kernel) at their largest UNet shapes, summarised
# with tools/ncu_summary.py.  Usage (under gpurun): bash tools/r2_ncu_k9_k10.sh
OUT=gpurun_out/r02_ncu_k9_k10; mkdir -p $OUT
timeout 300 ncu --set full --clock-control none -k regex:shortcut_tc_kernel -s 3 -c 2 -f -o $OUT/prof_k10 python tools/shortcut_probe.py 64 > $OUT/ncu_k10.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:linear_tc_kernel -s 3 -c 2 -f -o $OUT/prof_k9 python tools/linear_probe.py 64 > $OUT/ncu_k9.log 2>&1
python tools/ncu_summary.py $OUT/prof_k10.ncu-rep > $OUT/ncu_k10_summary.txt 2>&1
python tools/ncu_summary.py $OUT/prof_k9.ncu-rep > $OUT/ncu_k9_summary.txt 2>&1
for f in k10 k9; do ncu -i $OUT/prof_$f.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
seen=set()
for r in rows[1:]:
    d=dict(zip(h,r)); n=d.get('Metric Name','')
    if d.get('ID')!='0': continue
    if any(k in n for k in ['Duration','DRAM Throughput','L2 Cache Throughput','Shared Memory','Executed Ipc Active','Registers Per','Achieved Occ','L1/TEX Hit','Mem Busy','Mem Pipes']):
        print(d.get('Section Name','')[:28].ljust(28), n[:44].ljust(44), d.get('Metric Unit','')[:10].ljust(10), d.get('Metric Value',''))
" > $OUT/ncu_${f}_details.txt; done
cat $OUT/ncu_k10_summary.txt $OUT/ncu_k9_summary.txt | cut -c1-170
rm -f $OUT/*.ncu-rep
