mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/aa_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/aa_pytest_gpu.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/aa_pytest_gpu.txt | head -20
timeout 600 python -m pytest tests/test_gpu_generator.py tests/test_gpu_loop.py tests/test_gpu_video.py -q -s -k "bench_depth or decisive or oracle" 2>&1 | grep -E "rel-L2|pixels|frame|grad" | cut -c1-600
