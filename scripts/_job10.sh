set -x
mkdir -p gpurun_out; rm -f gpurun_out/j10_*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 300 python scripts/variant_select.py --install > gpurun_out/j10_variants.log 2>&1
tail -5 gpurun_out/j10_variants.log | cut -c1-330
BI=$(python -c "
import sys; sys.path.insert(0,'3dgsconverter_b200')
from gsx import _abi; print(_abi.lib.gsx_build_info().decode())")
echo "installed build: $BI"
timeout 800 python -u -m pytest tests -q -m gpu -p no:cacheprovider --timeout 200 --durations=5 > gpurun_out/j10_pytest.log 2>&1
tail -12 gpurun_out/j10_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sor_knn -c 1 -f -o gpurun_out/r02c_knn python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/j10_ncu_knn.log 2>&1
tail -2 gpurun_out/j10_ncu_knn.log
python scripts/ncu_kernel_summary.py gpurun_out/r02c_knn.ncu-rep k_sor_knn profiles/r02c_knn_ncu.json n=10000000 kind=mixed hash=i32wrap "build_info=$BI" > /dev/null 2> gpurun_out/j10_summary.err
cp profiles/r02c_knn_ncu.json gpurun_out/ 2>/dev/null
timeout 500 python bench.py > gpurun_out/j10_bench_n1.json 2> gpurun_out/j10_bench_n1.err
cut -c1-300 gpurun_out/j10_bench_n1.json; tail -3 gpurun_out/j10_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/j10_bench_n1.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'pageable',d['e2e']['pageable_input']['value'],'clocks',d.get('clocks'),'km e2e',d['kmeans'].get('e2e',{}).get('ms_per_call'))"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/j10_launches.log 2>&1
timeout 250 python scripts/copy_threads_probe.py > gpurun_out/j10_copy_probe.log 2>&1; cut -c1-700 gpurun_out/j10_copy_probe.log
