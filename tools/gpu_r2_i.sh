mkdir -p gpurun_out
timeout 600 python tools/bench_conv2.py --wgrad > gpurun_out/i_wgrad.txt 2>&1; echo "rc=$?" >> gpurun_out/i_wgrad.txt
timeout 600 python tools/bench_conv2.py --probe > gpurun_out/i_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/i_probe.txt
timeout 600 python tools/bench_conv2.py --time > gpurun_out/i_time.txt 2>&1
SPI_CONV_WGRAD=cudnn timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest_gpu_wgrad_cudnn.txt 2>&1; echo "rc=$?" >> gpurun_out/i_pytest_gpu_wgrad_cudnn.txt
SPI_CONV_WGRAD=cudnn timeout 600 python bench.py --no-cpu-baseline > gpurun_out/i_bench_wgrad_cudnn.json 2> gpurun_out/i_bench_wgrad_cudnn.err
SPI_CONV_ENGINE=cudnn timeout 600 python bench.py --no-cpu-baseline > gpurun_out/i_bench_cudnn.json 2> gpurun_out/i_bench_cudnn.err
cat gpurun_out/i_wgrad.txt; grep -v "e-04" gpurun_out/i_probe.txt; cat gpurun_out/i_time.txt; tail -15 gpurun_out/i_pytest_gpu_wgrad_cudnn.txt; head -c 400 gpurun_out/i_bench_wgrad_cudnn.json; echo; tail -3 gpurun_out/i_bench_wgrad_cudnn.err; head -c 400 gpurun_out/i_bench_cudnn.json
