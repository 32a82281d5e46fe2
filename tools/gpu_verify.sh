# Round-end verification on one B200: GPU tests, smoke, bench line, per-kernel device-time breakdown, renderer alone,
# ncu launch list and full captures of the two renderer kernels.
# Usage: gpurun --timeout 3000 -- "bash tools/gpu_verify.sh"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/smoke.txt
timeout 900 python bench.py --gpus 1 --steps 24 --warmup 6 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?" >> gpurun_out/bench_final.err
timeout 300 python tools/profile_step.py > gpurun_out/profile_eager.txt 2>&1
timeout 300 python tools/bench_render.py > gpurun_out/bench_render.txt 2>&1
timeout 300 python tools/bench_stream.py > gpurun_out/bench_stream.txt 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 30000 --csv --log-file gpurun_out/launches_all.csv python tools/ncu_target.py > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'render_fwd_tc|render_bwd_tc' -s 2 -c 2 -o gpurun_out/prof_render_tc_final python tools/bench_render.py > gpurun_out/ncu_render_final.log 2>&1
tail -4 gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/smoke.txt; cat gpurun_out/bench_final.json | head -c 600; echo; cat gpurun_out/bench_render.txt
