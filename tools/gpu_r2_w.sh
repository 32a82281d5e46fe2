mkdir -p gpurun_out
timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "rc=$?" >> gpurun_out/w_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/w_bench.json').read().strip().splitlines()[-1])
r=d['roofline']; print(d['value'])
for e in r['conv_engine']['by_shape_top']: print(e)
PY
tail -3 gpurun_out/w_bench.err
