mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 6 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?" >> gpurun_out/bench2.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_all.csv python tools/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
tail -5 gpurun_out/pytest_gpu.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
PY
wc -l gpurun_out/launches_all.csv
