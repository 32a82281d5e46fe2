mkdir -p gpurun_out
timeout 300 python tools/bench_conv2.py --wgrad-debug > gpurun_out/m_wgrad_debug.txt 2>&1
timeout 600 python tools/bench_conv2.py --wgrad > gpurun_out/m_wgrad.txt 2>&1; echo "rc=$?" >> gpurun_out/m_wgrad.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/m_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/m_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/m_bench_tc2.json 2> gpurun_out/m_bench_tc2.err
SPI_CONV_WGRAD=cudnn timeout 600 python bench.py --no-cpu-baseline > gpurun_out/m_bench_wgrad_cudnn.json 2> gpurun_out/m_bench_wgrad_cudnn.err
head -3 gpurun_out/m_wgrad_debug.txt; cat gpurun_out/m_wgrad.txt; tail -15 gpurun_out/m_pytest_gpu.txt; head -c 300 gpurun_out/m_bench_tc2.json; echo; tail -3 gpurun_out/m_bench_tc2.err; head -c 300 gpurun_out/m_bench_wgrad_cudnn.json; echo; tail -3 gpurun_out/m_bench_wgrad_cudnn.err
