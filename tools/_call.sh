mkdir -p gpurun_out
timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv.txt 2>&1; echo "rc=$?" >> gpurun_out/bench_conv.txt
cat gpurun_out/bench_conv.txt
