# Round-end evidence pass on one B200 (ordered by priority; every step bounded): GPU test suite, default bench line, smoke,
# launch list of the sampling graph, reference arm.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/final_pytest_gpu.txt
timeout 420 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.txt 2>&1
MDTB200_BN_D=192 MDTB200_BENCH_SKIP_EXTRAS=1 timeout 120 python bench.py --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null | cut -c1-300 > gpurun_out/final_bn_d192.txt
MDTB200_BENCH_SKIP_EXTRAS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 3000 -c 1500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv > gpurun_out/r02_launches_table.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
MDTB200_BRANCHES=1 timeout 120 python tools/ktrace.py 64 2 > gpurun_out/final_ktrace.txt 2>&1
tail -3 gpurun_out/final_pytest_gpu.txt; tail -c 600 gpurun_out/final_bench.json; tail -2 gpurun_out/final_smoke.txt; cat gpurun_out/final_bn_d192.txt | cut -c1-200
