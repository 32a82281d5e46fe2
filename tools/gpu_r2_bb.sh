mkdir -p gpurun_out
timeout 600 python bench.py --config 2 --steps 24 --warmup 6 > gpurun_out/bb_bench_cfg2.json 2> gpurun_out/bb_bench_cfg2.err; echo "rc=$?" >> gpurun_out/bb_bench_cfg2.err
timeout 600 python bench.py --config 4 --no-cpu-baseline > gpurun_out/bb_bench_cfg4_48.json 2> gpurun_out/bb_bench_cfg4_48.err; echo "rc=$?" >> gpurun_out/bb_bench_cfg4_48.err
timeout 600 python bench.py --config 4 --depth 96 96 --no-cpu-baseline > gpurun_out/bb_bench_cfg4_96.json 2> gpurun_out/bb_bench_cfg4_96.err; echo "rc=$?" >> gpurun_out/bb_bench_cfg4_96.err
timeout 600 python bench.py --config 4 --depth 48 48 --nrr 64 --no-cpu-baseline > gpurun_out/bb_bench_cfg4_48_nrr64.json 2> gpurun_out/bb_bench_cfg4_48_nrr64.err; echo "rc=$?" >> gpurun_out/bb_bench_cfg4_48_nrr64.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel|conv_wgrad_tc2_kernel" -c 9 -o gpurun_out/bb_prof_conv python tools/bench_conv2.py --ncu-target > gpurun_out/bb_ncu_conv.log 2>&1
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40000 --csv --log-file gpurun_out/bb_launches_all.csv python tools/ncu_target.py > gpurun_out/bb_ncu_launches.log 2>&1
for f in gpurun_out/bb_bench_cfg*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d['value'], d['e2e']['value'] if d.get('e2e') else None, d['config']['step_mix'], d['config']['depth_samples'], d['config']['neural_rendering_resolution'], d.get('cpu_baseline'))
except Exception as e: print('ERR', e)
PY
done
tail -2 gpurun_out/bb_bench_cfg2.err; tail -3 gpurun_out/bb_ncu_conv.log; tail -3 gpurun_out/bb_ncu_launches.log; ls -la gpurun_out/bb_*
