mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/t_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/t_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/t_bench_tc2.json 2> gpurun_out/t_bench_tc2.err
timeout 600 python tools/profile_step.py > gpurun_out/t_profile_tc2.txt 2>&1
grep -E "passed|failed|FAILED" gpurun_out/t_pytest_gpu.txt | head -20; head -c 300 gpurun_out/t_bench_tc2.json; echo; head -30 gpurun_out/prof_tc2_rot.txt
