#!/bin/bash
# measurement of a training-kernel change: training tests, step bench (both LSTM kernel versions), launch list, captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/pytest_train.log 2>&1; tail -4 gpurun_out/pytest_train.log
timeout 600 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 --cpu 0 > gpurun_out/train_bench.json 2> gpurun_out/train_bench.err; tail -c 900 gpurun_out/train_bench.json; tail -3 gpurun_out/train_bench.err
timeout 600 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 --cpu 0 --one-row 1 > gpurun_out/train_bench_one_row.json 2>> gpurun_out/train_bench.err; tail -c 900 gpurun_out/train_bench_one_row.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train.csv python tools/train_bench.py --batch 8 --seconds 5 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_train_bwd|lstm_train_fwd|outer_kernel|ln_bwd|rowgemm' -s 20 -c 22 -o gpurun_out/prof_train python tools/train_bench.py --batch 4 --seconds 2 --steps 1 --warmup 0 --cpu 0 > gpurun_out/ncu_t3.log 2>&1
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches_train.csv') if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',row['Kernel Name'])+" "+row['Grid Size']; v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v
    tot[name]+=v; cnt[name]+=1
T=sum(tot.values())
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:14]:
    print("%-75s n=%4d total %9.1f us avg %8.1f %5.1f%%"%(k[:75],cnt[k],v,v/cnt[k],100*v/T))
print("total",T)
PY
