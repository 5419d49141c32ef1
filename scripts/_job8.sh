set -x
mkdir -p gpurun_out; rm -f gpurun_out/j8_*
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
# 1. A/B of the query-kernel variants with parity digests; the fastest eligible one becomes lib/libgsx.so for the rest
timeout 500 python scripts/variant_select.py --install > gpurun_out/j8_variants.log 2>&1
tail -9 gpurun_out/j8_variants.log
BI=$(python -c "
import sys; sys.path.insert(0,'3dgsconverter_b200')
from gsx import _abi; print(_abi.lib.gsx_build_info().decode())")
echo "installed build: $BI"
# 2. the whole GPU suite on that build
timeout 700 python -u -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 200 --durations=8 > gpurun_out/j8_pytest.log 2>&1
tail -14 gpurun_out/j8_pytest.log
# 3. ncu capture of the query kernel inside bench.py, summarised on the box so that the bench line can use it
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sor_knn -c 1 -f -o gpurun_out/r02c_knn python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/j8_ncu_knn.log 2>&1
tail -2 gpurun_out/j8_ncu_knn.log
python scripts/ncu_kernel_summary.py gpurun_out/r02c_knn.ncu-rep k_sor_knn profiles/r02c_knn_ncu.json n=10000000 kind=mixed hash=i32wrap "build_info=$BI" > /dev/null 2> gpurun_out/j8_summary.err
cp profiles/r02c_knn_ncu.json gpurun_out/ 2>/dev/null
# 4. the bench line and the launch list
timeout 500 python bench.py > gpurun_out/j8_bench_n1.json 2> gpurun_out/j8_bench_n1.err
cut -c1-700 gpurun_out/j8_bench_n1.json; tail -3 gpurun_out/j8_bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/j8_launches.log 2>&1
