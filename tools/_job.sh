timeout 900 python -m pytest tests/test_gpu_train_ops2.py tests/test_gpu_training.py tests/test_gpu_optim.py tests/test_perceiver.py -m gpu -x -q 2>&1 | tail -5
python tools/train_once.py 2>&1 | grep -o "\"eager\": {\"value\": [0-9.]*, \"ms_per_step\": [0-9.]*\|\"graphed\": {\"value\": [0-9.]*, \"ms_per_step\": [0-9.]*\|Error.*"
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import bench
print(bench.perceiver_measure(torch.device("cuda"), 256))
PY
