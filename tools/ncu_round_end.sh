mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_final.csv python tools/ncu_target.py > gpurun_out/ncu_target.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_wgrad_tc2 -c 2 -o gpurun_out/r2_rows_outer python tools/bench_rows_outer.py > gpurun_out/ncu_rows.log 2>&1
echo "rows rc=$?"
ls -la gpurun_out | tail -5
