mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc2.py -x -q -s > gpurun_out/o_pytest_conv2.txt 2>&1; echo "rc=$?" >> gpurun_out/o_pytest_conv2.txt
timeout 600 python tools/bench_conv2.py --probe > gpurun_out/o_probe.txt 2>&1
timeout 600 python tools/bench_conv2.py --one > gpurun_out/o_one.txt 2>&1
timeout 600 python tools/bench_conv2.py --time > gpurun_out/o_time.txt 2>&1
timeout 600 python tools/bench_conv2.py --wgrad > gpurun_out/o_wgrad.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/o_bench_tc2.json 2> gpurun_out/o_bench_tc2.err
grep -E "passed|failed|rel-L2|tc2 vs|Error" gpurun_out/o_pytest_conv2.txt | head; grep -vc "e-04" gpurun_out/o_probe.txt; grep -v "e-04" gpurun_out/o_probe.txt | head; cat gpurun_out/o_one.txt gpurun_out/o_time.txt; grep "TF/s\|rel-L2 [^2]" gpurun_out/o_wgrad.txt; head -c 300 gpurun_out/o_bench_tc2.json; tail -2 gpurun_out/o_bench_tc2.err
