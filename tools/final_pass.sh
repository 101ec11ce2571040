# Round-end evidence pass on one B200: GPU test suite, default bench line, reference arm, launch lists.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/final_pytest_gpu.txt
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv python tools/train_once.py eager 3 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_train_launches.csv > gpurun_out/r02_train_summary.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_perceiver_launches.csv python tools/perceiver_once.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_perceiver_launches.csv > gpurun_out/r02_perceiver_summary.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
tail -3 gpurun_out/final_pytest_gpu.txt; tail -c 600 gpurun_out/final_bench.json; tail -2 gpurun_out/final_smoke.txt
