for m in 4 5; do
  TES_TC5_NPOLY=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cmp_m$m.csv python scripts/dev_screen_one.py 1000000 6 1024 1024 1 $m > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/cmp_m$m.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': h=r; st=i+2; break
ix={k:i for i,k in enumerate(h)}
seen={}
for r in rows[st:]:
    if len(r)<len(h): continue
    n=r[ix['Kernel Name']].split('(')[0][:40]; seen.setdefault(n,[]).append(float(r[ix['Metric Value']])/1e6)
print('method $m', {k: round(v[-1],3) for k,v in seen.items() if k.startswith('tes') or 'tes::' in k or 'rff' in k})
PY
done
