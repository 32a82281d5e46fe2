# Round-end verification on one B200: GPU tests, smoke, the driver's bench commands (both arms), conv-engine table.
# Usage: gpurun --timeout 3000 -- "bash tools/gpu_verify.sh"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/v_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/v_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.txt 2>&1; echo "rc=$?" >> gpurun_out/v_smoke.txt
( time timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/v_bench_n1.json 2> gpurun_out/v_bench_n1.err; echo "rc=$?" >> gpurun_out/v_bench_n1.err
( time timeout 900 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/v_bench_ref.json 2> gpurun_out/v_bench_ref.err; echo "rc=$?" >> gpurun_out/v_bench_ref.err
( timeout 600 python tools/bench_conv2.py --time; timeout 600 python tools/bench_conv2.py --wgrad; timeout 600 python tools/bench_conv2.py --small ) > gpurun_out/v_conv_engine.txt 2>&1
grep -E "passed|failed" gpurun_out/v_pytest_gpu.txt; tail -3 gpurun_out/v_smoke.txt; head -c 400 gpurun_out/v_bench_n1.json; echo; tail -4 gpurun_out/v_bench_n1.err; head -c 300 gpurun_out/v_bench_ref.json; echo; tail -4 gpurun_out/v_bench_ref.err
