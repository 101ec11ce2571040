python tools/demo_drop_in.py > gpurun_out/demo.log 2>&1; grep -v "Warning\|warn(\|^  " gpurun_out/demo.log | head -40 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -x -q 2>&1 | tail -3
