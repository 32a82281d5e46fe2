mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc2.py -q -s -k "repeatab" > gpurun_out/r_repeat.txt 2>&1
tail -30 gpurun_out/r_repeat.txt
