set -x
mkdir -p gpurun_out; rm -f gpurun_out/j6_*
timeout 300 python -u -m pytest tests/test_morton_gpu.py tests/test_dropin_api_gpu.py tests/test_records_gpu.py tests/test_sor_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 150 > gpurun_out/j6_pytest1.log 2>&1
tail -5 gpurun_out/j6_pytest1.log
timeout 400 python -u -m pytest tests/test_multigpu_nccl.py -x -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/j6_nccl.log 2>&1
tail -5 gpurun_out/j6_nccl.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/j6_bench_n2.json 2> gpurun_out/j6_bench_n2.err
cut -c1-300 gpurun_out/j6_bench_n2.json; tail -3 gpurun_out/j6_bench_n2.err
