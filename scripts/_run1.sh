timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mincut" 2>&1 | tail -4
