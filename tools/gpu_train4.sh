#!/bin/bash
# A/B of packed FMAs in the training GEMM kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/pytest_train.log 2>&1; tail -3 gpurun_out/pytest_train.log
timeout 200 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 --cpu 0 --ffma2 0 > gpurun_out/train_bench_ffma2_0.json 2> gpurun_out/train_bench.err; tail -c 500 gpurun_out/train_bench_ffma2_0.json
timeout 200 python tools/train_bench.py --batch 8 --seconds 5 --steps 3 --cpu 0 --ffma2 1 > gpurun_out/train_bench_ffma2_1.json 2>> gpurun_out/train_bench.err; tail -c 500 gpurun_out/train_bench_ffma2_1.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_train_ffma2.csv python tools/train_bench.py --batch 8 --seconds 5 --steps 1 --warmup 0 --cpu 0 --ffma2 1 > gpurun_out/ncu_t4.log 2>&1
python - <<'PY'
import csv, collections, re
lines=[l for l in open('gpurun_out/launches_train_ffma2.csv') if not l.startswith('==')]
tot=collections.defaultdict(float); cnt=collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',row['Kernel Name'])+" "+row['Grid Size']; v=float(row['Metric Value'].replace(',','')); u=row['Metric Unit']
    v = v/1e3 if u=='ns' else v*1e3 if u=='ms' else v
    tot[name]+=v; cnt[name]+=1
T=sum(tot.values())
for k,v in sorted(tot.items(), key=lambda x:-x[1])[:9]:
    print("%-75s n=%4d total %9.1f us avg %8.1f %5.1f%%"%(k[:75],cnt[k],v,v/cnt[k],100*v/T))
print("total",T)
PY
