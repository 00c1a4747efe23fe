python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "digest_sort or candidate_first or search_at_load or default_search or staged or enzymes or isoleucine" 2>&1 | tail -5
run() { python bench.py --no-cpu --e2e-steps 5 --steps 3 --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.4g  ms/step %.2f  e2e %.2f'%(d['value'], d['ms_per_step'], d['e2e']['ms_per_step']))"; }
run msd
python tools/trace_e2e.py 2>&1 | grep -v "k0_prepare\|gather_peaks\|k_wait_chunk" | tail -42
