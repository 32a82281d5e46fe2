# Round-end verification on one B200: GPU tests, smoke, bench line, per-kernel device-time breakdown, renderer alone.
# Usage: gpurun --timeout 2400 -- "bash tools/gpu_verify.sh"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 6 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?" >> gpurun_out/bench_final.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
timeout 300 python tools/bench_render.py > gpurun_out/bench_render.txt 2>&1
tail -4 gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/smoke.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
r=d['roofline']
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], r['ms_per_image'], r['frac'], r['render_fwd_tc_kernel']['ms_per_image'], d['cpu_baseline']['value'])
PY
cat gpurun_out/bench_render.txt
