set -x
mkdir -p gpurun_out; rm -f gpurun_out/j3_*
timeout 600 python -u -m pytest tests/test_sor_gpu.py tests/test_goldens_gpu.py tests/test_radix_sort_gpu.py tests/test_dropin_api_gpu.py tests/test_stats_dist_gpu.py tests/test_kmeans_density_masks_gpu.py tests/test_kmeans_prefilter_gpu.py tests/test_kmeans_tc_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 150 --durations=12 > gpurun_out/j3_pytest.log 2>&1
tail -25 gpurun_out/j3_pytest.log
for v in k16mb4 k16mb5 k16mb6 k16mb8 k16t3 k16t6 warp; do
  echo "== $v" >> gpurun_out/j3_knn_variants.log
  GSX_LIB=$PWD/3dgsconverter_b200/lib/variants/libgsx_$v.so timeout 120 python scripts/sor_probe.py 10000000 mixed,uniform >> gpurun_out/j3_knn_variants.log 2>&1
done
grep -E "==|i32wrap" gpurun_out/j3_knn_variants.log | cut -c1-200
timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j3_stream.json 2> gpurun_out/j3_stream.err
python -c "
import json;d=json.load(open('gpurun_out/j3_stream.json'))
for k,v in d['stages'].items(): print(k, v['ms'], v['frac_of_hbm_peak'])"
