#!/bin/bash
mkdir -p gpurun_out
for c in serial4 div2; do timeout 300 python benchmarks/bookkeeping.py --config $c --envs $([ $c = serial4 ] && echo 65536 || echo 262144); done > gpurun_out/r2_bookkeeping.jsonl 2> gpurun_out/r2_bookkeeping.err
cat gpurun_out/r2_bookkeeping.jsonl; tail -3 gpurun_out/r2_bookkeeping.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches_book.csv python benchmarks/bookkeeping.py > gpurun_out/r2_ncu_book.log 2>&1
grep -v "^==" gpurun_out/r2_launches_book.csv | python -c "
import csv,sys,collections
rows=list(csv.DictReader(sys.stdin)); agg=collections.defaultdict(list)
for r in rows: agg[r['Kernel Name'][:60]].append(float(r['Metric Value'].replace(',','')))
for k,v in agg.items(): print(len(v), round(sum(v)/len(v)/1000,2), k)
"
