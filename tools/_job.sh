timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cabi.py tests/test_r2_pins.py -m gpu -x -q 2>&1 | tail -4
