timeout 900 python -m pytest tests/test_gpu_train_ops2.py tests/test_gpu_training.py tests/test_gpu_optim.py -m gpu -x -q 2>&1 | tail -3
python tools/train_once.py 2>&1 | grep -o "\"eager\": {\"value\": [0-9.]*, \"ms_per_step\": [0-9.]*\|\"graphed\": {\"value\": [0-9.]*, \"ms_per_step\": [0-9.]*\|Error.*"
