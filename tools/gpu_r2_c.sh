mkdir -p gpurun_out
timeout 600 python tools/bench_conv2.py --probe > gpurun_out/c_conv2_probe.txt 2>&1; echo "rc=$?" >> gpurun_out/c_conv2_probe.txt
timeout 600 python tools/bench_conv2.py --time > gpurun_out/c_conv2_time.txt 2>&1; echo "rc=$?" >> gpurun_out/c_conv2_time.txt
timeout 900 python -m pytest tests/test_gpu_loop.py -x -q -k "graph_replay or graphed_projector" -s > gpurun_out/c_pytest_graph.txt 2>&1; echo "rc=$?" >> gpurun_out/c_pytest_graph.txt
grep -c "e-04" gpurun_out/c_conv2_probe.txt; grep -v "e-04" gpurun_out/c_conv2_probe.txt; cat gpurun_out/c_conv2_time.txt; grep "iteration\|w_opt\|passed\|failed" gpurun_out/c_pytest_graph.txt
