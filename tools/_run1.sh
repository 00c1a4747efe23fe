python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ladder or default_search or golden" 2>&1 | tail -3
run() { python bench.py --no-cpu --e2e-steps 0 --steps 3 --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.4g  ms/step %.2f step3 %.2f'%(d['value'], d['ms_per_step'], d['kernel_ms_per_step']['k1_step3']))"; }
run b64s1
PQ_K1_STAGES=2 run b64s2
PQ_STEP3_BLOCK=128 run b128s1
PQ_STEP3_BLOCK=128 PQ_K1_STAGES=2 run b128s2
PQ_STEP3_BLOCK=128 PQ_K1_STAGES=2 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ladder or default_search or golden" 2>&1 | tail -3
