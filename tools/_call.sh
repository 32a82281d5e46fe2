mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 6 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
timeout 300 python tools/bench_stream.py > gpurun_out/bench_stream.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "timed/" -c 9000 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'render_fwd_tc|render_bwd_tc' -s 2 -c 4 -o gpurun_out/prof_render_tc_final python tools/bench_render.py > gpurun_out/ncu_render_final.log 2>&1
tail -3 gpurun_out/pytest_gpu.txt; cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err; wc -l gpurun_out/launches.csv
