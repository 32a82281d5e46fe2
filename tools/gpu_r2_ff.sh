mkdir -p gpurun_out
timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/ff_bench.json 2> gpurun_out/ff_bench.err; echo "rc=$?" >> gpurun_out/ff_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/ff_bench.json').read().strip().splitlines()[-1])
r=d['roofline']; print(d['value'])
for t in ('bias_act','upfirdn2d'):
    print(t, {k:v for k,v in r['streaming_kernels'][t].items() if k!='by_shape_top'})
    for e in r['streaming_kernels'][t]['by_shape_top']: print('   ', e)
PY
tail -2 gpurun_out/ff_bench.err | cut -c1-200
timeout 600 python bench.py --config 2 --steps 24 --warmup 6 > gpurun_out/ff_bench_cfg2.json 2> gpurun_out/ff_bench_cfg2.err; echo "cfg2 rc=$?"; head -c 400 gpurun_out/ff_bench_cfg2.json; grep -v "^frame\|Warning\|warn" gpurun_out/ff_bench_cfg2.err | tail -5
