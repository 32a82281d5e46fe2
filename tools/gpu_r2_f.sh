mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__cycles_active.avg --clock-control none -k regex:"conv_tc2|conv_fprop_tc05|cutlass|cudnn|gemm|xmma" -c 120 --csv --log-file gpurun_out/f_launches.csv python tools/bench_conv2.py --time3 > gpurun_out/f_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/f_launches.csv')))
i0=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[i0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
seq=[]
for r in rows[i0+2:]:
    if len(r)>vi and r[mi]=='gpu__time_duration.sum':
        seq.append((r[ki][:60], r[vi]))
for k,v in seq[:90]: print(k,v)
PY
