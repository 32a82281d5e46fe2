timeout 600 python -m pytest tests/test_gpu_render.py tests/test_legacy_pkl.py -x -q 2>&1 | tail -3
python -c "
from spi_b200 import _lib
print('tc error flag:', _lib.load().spi_tc_error())"
timeout 300 python tools/bench_render.py 2>&1 | grep "planes only"
