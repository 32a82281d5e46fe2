# Round 2, call A: the driver's exact bench commands (both arms) at N=1, then the GPU test-suite.
mkdir -p gpurun_out
nproc > gpurun_out/a_nproc.txt
( time timeout 900 python3 bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/a_bench_ours_s20.json 2> gpurun_out/a_bench_ours_s20.err; echo "rc=$?" >> gpurun_out/a_bench_ours_s20.err
( time timeout 900 python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/a_bench_ref_s20.json 2> gpurun_out/a_bench_ref_s20.err; echo "rc=$?" >> gpurun_out/a_bench_ref_s20.err
( time timeout 600 python3 bench.py --no-cpu-baseline ) > gpurun_out/a_bench_ours_default.json 2> gpurun_out/a_bench_ours_default.err; echo "rc=$?" >> gpurun_out/a_bench_ours_default.err
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/a_pytest_gpu.txt
tail -5 gpurun_out/a_pytest_gpu.txt; head -c 1500 gpurun_out/a_bench_ours_s20.json; echo; tail -5 gpurun_out/a_bench_ours_s20.err; head -c 800 gpurun_out/a_bench_ref_s20.json; tail -5 gpurun_out/a_bench_ref_s20.err
