set -x
mkdir -p gpurun_out; rm -f gpurun_out/j11_*
nvidia-smi --query-gpu=name --format=csv,noheader
timeout 400 python -u -m pytest tests/test_multigpu_nccl.py tests/test_hostcopy_gpu.py tests/test_reference_kernels_pin.py -q -m gpu -p no:cacheprovider --timeout 300 > gpurun_out/j11_pytest.log 2>&1
tail -6 gpurun_out/j11_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/j11_bench_n2.json 2> gpurun_out/j11_bench_n2.err
cut -c1-400 gpurun_out/j11_bench_n2.json; tail -3 gpurun_out/j11_bench_n2.err
python -c "
import json; d=json.load(open('gpurun_out/j11_bench_n2.json'))
print('value',d['value'],d['ms_per_step'],'clocks',d.get('clocks'),'parity',d.get('parity'),'stage',d.get('stage_ms'))"
