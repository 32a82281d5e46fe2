mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boxcx.py tests/test_gpu_conv_tc2.py tests/test_gpu_loop.py -q > gpurun_out/u_pytest.txt 2>&1; echo "rc=$?" >> gpurun_out/u_pytest.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
grep -E "passed|failed|FAILED|Error" gpurun_out/u_pytest.txt | head -20; head -c 300 gpurun_out/u_bench.json
