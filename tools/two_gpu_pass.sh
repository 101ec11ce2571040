# 2-GPU evidence: DDP gradient test, sampling bench (weak scaling, no collective), DDP training workload.  Outputs under gpurun_out/.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r02_2gpu_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r02_2gpu_bench.json 2> gpurun_out/r02_2gpu_bench.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --workload train --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_2gpu_train.json 2> gpurun_out/r02_2gpu_train.err
timeout 600 python bench.py --workload train --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_1gpu_train.json 2> gpurun_out/r02_1gpu_train.err
cat gpurun_out/r02_2gpu_pytest.txt; head -c 700 gpurun_out/r02_2gpu_bench.json; echo; head -c 1200 gpurun_out/r02_2gpu_train.json; echo; head -c 1200 gpurun_out/r02_1gpu_train.json; tail -3 gpurun_out/r02_2gpu_train.err
