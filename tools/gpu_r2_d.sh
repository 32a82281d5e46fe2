mkdir -p gpurun_out
timeout 300 python tools/bench_conv2.py --one > gpurun_out/d_conv2_one.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc2_kernel -s 3 -c 1 -o gpurun_out/d_prof_conv_tc2 python tools/bench_conv2.py --one > gpurun_out/d_ncu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_loop.py -x -q -k "graph_replay or graphed_projector" -s > gpurun_out/d_pytest_graph.txt 2>&1; echo "rc=$?" >> gpurun_out/d_pytest_graph.txt
cat gpurun_out/d_conv2_one.txt; tail -3 gpurun_out/d_ncu.log; grep "iteration\|w_opt\|passed\|failed" gpurun_out/d_pytest_graph.txt
