mkdir -p gpurun_out
( time timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "rc=$?" >> gpurun_out/v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[-1])
r=d['roofline']
print(d['value'], d['e2e']['value'], d['gpu_launches'], d['cpu_baseline'])
print({k:(v if not isinstance(v,dict) else '...') for k,v in r.items()})
print(r.get('conv_engine')); print(r.get('streaming_kernels'))
rb=r.get('render_backward', r); print(rb['ms_per_image'], rb['render_fwd']['ms_per_image'], rb['frac'], rb['share_of_step_time'])
PY
tail -5 gpurun_out/v_bench.err
