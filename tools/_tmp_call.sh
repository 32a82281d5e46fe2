timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 6 --no-cpu-baseline > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench5.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
PY
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
