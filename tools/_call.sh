mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python tools/bench_stream.py > gpurun_out/bench_stream.txt 2>&1
timeout 600 python bench.py --gpus 1 --steps 24 --warmup 6 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'blur4|upfirdn2d_cfast' -c 4 -o gpurun_out/prof_blur python tools/bench_stream.py --quick > gpurun_out/ncu_blur.log 2>&1
tail -4 gpurun_out/pytest_gpu.txt; grep -E "epilogue_grad|blur4" gpurun_out/bench_stream.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['ms_per_image'], d['roofline']['streaming_kernels'])
PY
grep wall gpurun_out/profile_eager.txt
