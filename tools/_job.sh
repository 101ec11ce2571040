python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_perceiver.py -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import torch, sys
sys.path.insert(0, ".")
import bench
print(bench.perceiver_measure(torch.device("cuda"), 256))
PY
