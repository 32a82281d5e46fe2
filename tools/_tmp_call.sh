timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 24 --warmup 6 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?" >> gpurun_out/bench_n2.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e']['value'], d['clocks'])
PY
tail -1 gpurun_out/bench_n2.err
