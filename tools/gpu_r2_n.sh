mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc2.py -x -q -s > gpurun_out/n_pytest_conv2.txt 2>&1; echo "rc=$?" >> gpurun_out/n_pytest_conv2.txt
timeout 600 python tools/profile_step.py > gpurun_out/n_profile_tc2.txt 2>&1
SPI_CONV_ENGINE=cudnn timeout 600 python tools/profile_step.py > gpurun_out/n_profile_cudnn.txt 2>&1
grep -E "passed|failed|rel-L2|tc2 vs|Error|error" gpurun_out/n_pytest_conv2.txt | head -30
head -40 gpurun_out/prof_tc2_rot.txt; echo; head -40 gpurun_out/prof_cudnn_rot.txt
