mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc2.py -q -s -k "do_not_depend" > gpurun_out/q_engine_cmp.txt 2>&1
SPI_CONV_ENGINE=cudnn timeout 600 python -m pytest tests/test_gpu_generator.py -q -s -k "synthesis_gradients" > gpurun_out/q_grad_cudnn.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_generator.py -q -s -k "synthesis_gradients" > gpurun_out/q_grad_tc2.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_generator.py -q -s -k "synthesis_gradients" > gpurun_out/q_grad_tc2_b.txt 2>&1
grep "tc2 vs" gpurun_out/q_engine_cmp.txt; grep "^grad rel-L2" gpurun_out/q_grad_cudnn.txt gpurun_out/q_grad_tc2.txt gpurun_out/q_grad_tc2_b.txt
