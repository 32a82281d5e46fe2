mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_render.py -x -q 2>&1 | tail -2
timeout 300 python tools/bench_render.py 2>&1 | grep "32+32"
