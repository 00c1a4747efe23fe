run() { python bench.py --no-cpu --e2e-steps 5 --steps 2 --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 e2e %.2f'%(d['e2e']['ms_per_step']))"; }
for c in 2 3 4 6 8; do PQ_CHUNKS_PER_PART=$c run cpp$c; done
