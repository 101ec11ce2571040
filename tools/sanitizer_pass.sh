# compute-sanitizer over the round-2 kernels (memcheck everywhere, racecheck on the shared-memory heavy ones); logs under gpurun_out/
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
run() { name=$1; shift; timeout 900 $S "$@" > gpurun_out/sanitizer_r02_$name.log 2>&1; echo "== $name: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_r02_$name.log | tail -1) | $(grep -E 'passed|failed|smoke OK' gpurun_out/sanitizer_r02_$name.log | tail -1)"; }
run smoke --tool memcheck python -c "import __graft_entry__ as g; g.smoke()"
run train_ops --tool memcheck python -m pytest tests/test_gpu_train_ops2.py -m gpu -x -q -k "not 5120-1536"
run train_step --tool memcheck python -m pytest tests/test_gpu_training.py -m gpu -x -q -k "loss_and_all_gradients or input_gradients or shipped_dropout"
run perceiver --tool memcheck python -m pytest tests/test_perceiver.py -m gpu -x -q -k "small or mask"
run race_gemm --tool racecheck python -m pytest tests/test_gpu_train_ops2.py -m gpu -x -q -k "gemm16_three_roles and 200-128-64 or gelu16 or dgrad_with_gelu and 200"
run race_simt --tool racecheck python -m pytest tests/test_gpu_train_ops2.py -m gpu -x -q -k "ln_fwd16 or res_drop or attention_backward or narrow"
run race_perceiver --tool racecheck python -m pytest tests/test_perceiver.py -m gpu -x -q -k "small"
