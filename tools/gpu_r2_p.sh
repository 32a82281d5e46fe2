mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/p_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/p_pytest_gpu.txt
grep -E "passed|failed|FAILED|Error" gpurun_out/p_pytest_gpu.txt | head -40
