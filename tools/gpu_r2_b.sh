mkdir -p gpurun_out
timeout 600 python tools/bench_conv2.py --probe > gpurun_out/b_conv2_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/b_conv2_probe.txt
timeout 600 python tools/bench_conv2.py --time > gpurun_out/b_conv2_time.txt 2>&1; echo "rc=$?" >> gpurun_out/b_conv2_time.txt
timeout 900 python -m pytest tests/test_gpu_loop.py -x -q -k "graph_replay or graphed_projector" -s > gpurun_out/b_pytest_graph.txt 2>&1; echo "rc=$?" >> gpurun_out/b_pytest_graph.txt
cat gpurun_out/b_conv2_probe.txt; cat gpurun_out/b_conv2_time.txt; tail -15 gpurun_out/b_pytest_graph.txt
