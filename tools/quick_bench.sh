#!/bin/bash
# quick GPU check: a parity subset + the resident-input bench line (kernel times per step)
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "default_search or parameter_variations or golden or edge_case or degenerate" 2>&1 | tail -2
python bench.py --no-cpu --e2e-steps 0 --steps 3 "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.1f  frac %.3f'%(d['value'], d['ms_per_step'], d['roofline']['frac'])); print({k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
