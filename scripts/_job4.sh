set -x
mkdir -p gpurun_out; rm -f gpurun_out/j4_*
export GSX_LIB=$PWD/3dgsconverter_b200/lib/variants/libgsx_k16v2.so
timeout 120 python scripts/sor_probe.py 10000000 mixed,uniform > gpurun_out/j4_probe.log 2>&1
grep i32wrap gpurun_out/j4_probe.log | cut -c1-160
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_sor_knn16 -c 1 -f -o gpurun_out/r02_knn16 python scripts/sor_probe.py 10000000 uniform > gpurun_out/j4_ncu.log 2>&1
tail -3 gpurun_out/j4_ncu.log
unset GSX_LIB
timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j4_stream.json 2> gpurun_out/j4_stream.err
python -c "
import json;d=json.load(open('gpurun_out/j4_stream.json'))
for k,v in d['stages'].items(): print(k, v['ms'], v['frac_of_hbm_peak'])"
timeout 300 python -u -m pytest tests/test_dropin_api_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 150 2>&1 | tail -3
