#!/bin/bash
# in_apply grid size sweep (blocks per SM) under ncu: per-kernel time of one forward.   tools/gpu_inapply.sh <tag>
tag=$1; mkdir -p gpurun_out; out=gpurun_out/inapply_$tag.txt; : > $out
for bps in 8 16 32 64; do
  RIB_INAPPLY_BPS=$bps ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:in_apply --csv \
     --log-file gpurun_out/inapply_${tag}_$bps.csv python tools/profile_forward.py --batch 32 --size 512 --iters 1 > /dev/null 2>&1
  python - $bps gpurun_out/inapply_${tag}_$bps.csv >> $out <<'PY'
import csv,sys,collections
rows=[r for r in csv.reader(l for l in open(sys.argv[2]) if l.startswith('"'))]
ix={n:i for i,n in enumerate(rows[0])}
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ix['Kernel Name']].split('(')[0]; v=float(r[ix['Metric Value']].replace(',','')); u=r[ix['Metric Unit']]
    v=v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=v
print('blocks per SM',sys.argv[1],' '.join('%s: %d launches %.1f us'%(k,a[0],a[1]) for k,a in agg.items()))
PY
done
cat $out
