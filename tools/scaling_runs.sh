# usage: gpurun --gpus N -- bash tools/scaling_runs.sh N TAG  -> weak and strong scaling lines as the driver launches them
N=$1; TAG=$2
for mode in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --scaling $mode --no-configs --no-cpu > gpurun_out/bench_n${N}_${mode}_$TAG.json 2> gpurun_out/bench_n${N}_${mode}_$TAG.err; echo "$mode rc=$?"
  python - <<PY
import json
for l in open('gpurun_out/bench_n${N}_${mode}_$TAG.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('N=$N $mode value %.4g ms/step %.2f e2e %.4g (%.2f ms) per-rank %s pairs %.4g'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['per_rank_ms_per_step'], d['config']['pairs_per_step']))
PY
done
