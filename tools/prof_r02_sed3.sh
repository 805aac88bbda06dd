cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"sed" -c 40 --csv --log-file gpurun_out/r02_sed_launches.csv python tools/sed_bench.py 256 > gpurun_out/r02_sed_ncu2.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_sed_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); mi=H.index('Metric Name'); vi=H.index('Metric Value'); ii=H.index('ID')
out={}
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    out.setdefault((int(r[ii]), r[ki].split('(')[0].split('::')[-1]),{})[r[mi]]=r[vi]
for (i,k),m in sorted(out.items())[-14:]:
    print(i,k,' '.join('%s=%s'%(a.split('__')[1][:22],b) for a,b in m.items()))
PY
