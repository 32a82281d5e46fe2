mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt
python -c "
from spi_b200 import _lib
print('tc error flag:', _lib.load().spi_tc_error())" >> gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --gpus 1 --steps 24 --warmup 6 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
tail -4 gpurun_out/pytest_gpu.txt; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['ms_per_image'], d['roofline']['render_bwd_incl_decoder_grad_gemms_ms_per_image_eager'])
PY
tail -3 gpurun_out/bench.err
