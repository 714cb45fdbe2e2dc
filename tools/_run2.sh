for i in 1 2; do timeout 300 python tools/profile_train.py --iters 5 2>&1 | grep -E "forward\(save\)|backward" | tr '\n' ' '; echo; done
echo "BWD_RES=0"; TGGCN_BWD_RES=0 timeout 300 python tools/profile_train.py --iters 5 2>&1 | grep -E "forward\(save\)|backward" | tr '\n' ' '; echo
echo "GEMM16=0"; TGGCN_GEMM16=0 timeout 300 python tools/profile_train.py --iters 5 2>&1 | grep -E "forward\(save\)|backward" | tr '\n' ' '; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s22_train.csv python tools/profile_train.py --iters 1 > /dev/null 2>&1
python - <<PY
import csv, collections, re
lines=[l for l in open('gpurun_out/s22_train.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
# last third = the timed iteration; print kernels between geo_gcn_kernel and heads_kernel of the last forward
names=[re.sub(r'\(.*','',r['Kernel Name'])[:50] for r in rows]
vals=[float(r['Metric Value'].replace(',',''))/1000 for r in rows]
idx=[i for i,n in enumerate(names) if 'geo_bn_stats' in n or ('geo_gcn_kernel' in n)]
start=idx[-2] if len(idx)>=2 else idx[-1]
end=[i for i,n in enumerate(names) if 'heads_kernel' in n][-1]
tot=0
for i in range(start,end+1):
    print(f'{names[i]:52s} {vals[i]:9.1f}'); tot+=vals[i]
print('forward total us', tot)
PY
