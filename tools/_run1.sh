timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pageable or candidate_first or search_at_load or pinned_caller or without_a_charge" 2>&1 | tail -5
timeout 300 python bench.py --no-cpu --e2e-steps 5 --steps 2 --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('e2e %.2f  pageable %s'%(e['ms_per_step'], e['pageable']))"
