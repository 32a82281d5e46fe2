mkdir -p gpurun_out
for v in "SPI_CONV_FUSE=0" "SPI_TC2_FLAGS=128" "SPI_CONV_WGRAD=cudnn" "SPI_TC2_FLAGS=4"; do
  env $v timeout 600 python -m pytest tests/test_gpu_generator.py -q -s -k "synthesis_gradients" 2>&1 | grep "^grad rel-L2" | sed "s/^/$v /" | grep -o "^[A-Z_0-9=]* \|noise_strength': [0-9.e-]*"
done
