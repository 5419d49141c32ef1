set -x
mkdir -p gpurun_out; rm -f gpurun_out/j9_*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
timeout 500 python scripts/variant_select.py --install > gpurun_out/j9_variants.log 2>&1
tail -8 gpurun_out/j9_variants.log | cut -c1-400
BI=$(python -c "
import sys; sys.path.insert(0,'3dgsconverter_b200')
from gsx import _abi; print(_abi.lib.gsx_build_info().decode())")
echo "installed build: $BI"
timeout 800 python -u -m pytest tests -q -m gpu -p no:cacheprovider --timeout 200 --durations=8 > gpurun_out/j9_pytest.log 2>&1
tail -25 gpurun_out/j9_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sor_knn -c 1 -f -o gpurun_out/r02c_knn python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/j9_ncu_knn.log 2>&1
tail -2 gpurun_out/j9_ncu_knn.log
python scripts/ncu_kernel_summary.py gpurun_out/r02c_knn.ncu-rep k_sor_knn profiles/r02c_knn_ncu.json n=10000000 kind=mixed hash=i32wrap "build_info=$BI" > /dev/null 2> gpurun_out/j9_summary.err
cp profiles/r02c_knn_ncu.json gpurun_out/ 2>/dev/null
timeout 500 python bench.py > gpurun_out/j9_bench_n1.json 2> gpurun_out/j9_bench_n1.err
cut -c1-500 gpurun_out/j9_bench_n1.json; tail -3 gpurun_out/j9_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/j9_launches.log 2>&1
timeout 200 python scripts/copy_threads_probe.py > gpurun_out/j9_copy_probe.log 2>&1; tail -5 gpurun_out/j9_copy_probe.log
timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j9_stream.json 2> gpurun_out/j9_stream.err
GSX_DENSITY_BITMAP=0 timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j9_stream_hashset.json 2> gpurun_out/j9_stream_hashset.err
python -c "
import json
for f in ('j9_stream','j9_stream_hashset'):
    d=json.load(open('gpurun_out/'+f+'.json'))
    print(f, {k:(v['ms'],v['frac_of_hbm_peak']) for k,v in d['stages'].items() if 'member' in k or 'compact' in k})"
