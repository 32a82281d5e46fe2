mkdir -p gpurun_out
timeout 600 python tools/bench_conv2.py --time > gpurun_out/e_conv2_time.txt 2>&1; echo "rc=$?" >> gpurun_out/e_conv2_time.txt
timeout 300 python tools/bench_conv2.py --one > gpurun_out/e_conv2_one.txt 2>&1
cat gpurun_out/e_conv2_time.txt gpurun_out/e_conv2_one.txt
