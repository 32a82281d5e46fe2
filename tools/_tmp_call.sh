timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 6 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench6.json'))
r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], r['ms_per_image'], r['render_fwd_tc_kernel']['ms_per_image'], r['decoder_grad_gemms_ms_per_image_eager'])
PY
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
