python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cell_lists or default_search or golden or random_config" 2>&1 | tail -3
run() { python bench.py --no-cpu --e2e-steps 5 --steps 3 --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.4g  ms/step %.2f e2e %.2f'%(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'])); print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"; }
run k2lists
