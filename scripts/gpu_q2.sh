python -m pytest tests/test_gpu_parity.py -x -q -k "large_64bit" 2>&1 | tail -3
