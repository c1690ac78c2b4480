#!/bin/bash
# L2 reduction-unit counters of the scatter-add kernels (ours and the reference's JIT kernel)
mkdir -p gpurun_out
ncu --query-metrics > gpurun_out/ncu_metrics_all.txt 2>&1
WANT="lts__t_sectors_op_red lts__t_sectors_op_atom lts__t_requests_srcunit_tex_op_red lts__t_sectors_srcunit_tex_op_red lts__t_requests_op_red lts__t_sector_op_red_hit_rate lts__t_sectors_srcunit_tex_op_read lts__d_sectors_fill_device lts__t_sectors lts__t_requests"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct"
for w in $WANT; do
  if grep -q "^$w " gpurun_out/ncu_metrics_all.txt; then M="$M,$w.sum"; fi
done
echo "metrics: $M" | tee gpurun_out/scatter_l2.log
grep -E "^lts__t_(sectors|requests)[a-z_]*op_(red|atom)" gpurun_out/ncu_metrics_all.txt | awk '{print $1}' | head -40 >> gpurun_out/scatter_l2.log
timeout 600 ncu --metrics "$M" --clock-control none --csv --log-file gpurun_out/scatter_l2.csv python tools/scatter_l2_counters.py >> gpurun_out/scatter_l2.log 2>&1
echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/scatter_l2.csv') if l.startswith('"'))]
h=rows[0]; out={}
for r in rows[1:]:
    d=dict(zip(h,r))
    k=(d['ID'],d['Kernel Name'][:70])
    out.setdefault(k,{})[d['Metric Name']]=d['Metric Value']
for (i,k),m in out.items():
    if 'scatter' in k or 'drjit' in k.lower() or 'enoki' in k.lower():
        print(i,k); print('   ',{a.replace('lts__t_','').replace('.sum',''):b for a,b in m.items()})
PY
