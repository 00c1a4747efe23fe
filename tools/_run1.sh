python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "no_enzyme or enzymes or digest_sort or default_search or abi" 2>&1 | tail -5
