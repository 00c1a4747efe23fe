timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_r2k.json 2> gpurun_out/bench_n8_r2k.err; echo "rc=$?"; tail -c 300 gpurun_out/bench_n8_r2k.err; python - <<'PY'
import json
for l in open('gpurun_out/bench_n8_r2k.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('N=8 value %.4g ms/step %.2f e2e %.4g (%.2f ms) per-rank %s'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['per_rank_ms_per_step'])); print({k:(round(v.get('value',0)/1e9,2), v.get('ms_per_step')) for k,v in (d['configs'] or {}).items() if isinstance(v,dict)})
PY
nproc; free -g | head -2
