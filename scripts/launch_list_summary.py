import csv, collections, json, sys
src, benchjson, dst = sys.argv[1], sys.argv[2], sys.argv[3]
rows=list(csv.reader(open(src)))
hi=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
h=rows[hi]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); im=h.index('Metric Name'); iu=h.index('Metric Unit')
acc=collections.OrderedDict(); tot=0.0; n=0
for r in rows[hi+1:]:
    if len(r)<=iv or r[im]!='gpu__time_duration.sum': continue
    name=r[ik].split('(')[0].replace('void ','').replace('fhc::','')
    v=float(r[iv].replace(',','')); u=r[iu]
    ms = v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    a=acc.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms; tot+=ms; n+=1
out=['ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pval_|hist_distance|bh_|radix_|fill_f64|lbeta_|outlier|digest|gather_ne|mid_range|scatter" -c 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --extras ""',
     "(B200, 300 M contacts in file order, 1 GPU; every launch of a library kernel in the command: warm-up passes, 2 timed passes, the digest and the e2e leg, whose passes launch K3 slice by slice; cold-cache serialised times: compare SHARES)",
     "launches %d, total %.3f ms" % (n, tot)]
for k in sorted(acc, key=lambda k:-acc[k][1]):
    out.append("%-40s launches %4d total_ms %9.3f share %5.1f%%" % (k[:40], acc[k][0], acc[k][1], 100*acc[k][1]/tot))
d=json.load(open(benchjson)); ks=d['kernels']; kt=sum(v['ms_per_step'] for v in ks.values())
out.append("")
out.append("bench.py (CUDA events inside the library, same workload): "+", ".join("%s %.2f ms = %.0f %%" % (k.replace('_kernel',''), v['ms_per_step'], 100*v['ms_per_step']/kt) for k,v in list(ks.items())[:7])+" of %.1f ms kernel time per step" % kt)
open(dst,'w').write("\n".join(out)+"\n"); print("\n".join(out))
