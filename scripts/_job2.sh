set -x
mkdir -p gpurun_out
rm -f gpurun_out/j2_*
for v in k16mb4 k16mb5 k16mb6 k16t4 k16t7 warp; do
  echo "== $v" >> gpurun_out/j2_knn_variants.log
  GSX_LIB=$PWD/3dgsconverter_b200/lib/variants/libgsx_$v.so timeout 200 python scripts/sor_probe.py 10000000 mixed,uniform >> gpurun_out/j2_knn_variants.log 2>&1
done
for v in onesweep threek; do
  GSX_LIB=$PWD/3dgsconverter_b200/lib/variants/libgsx_$v.so timeout 300 python scripts/sort_ab.py $v >> gpurun_out/j2_sort_ab.log 2>&1
done
timeout 200 python scripts/build_profile.py 80000000 > gpurun_out/j2_build80.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/j2_build80_launches.csv python scripts/build_profile.py 80000000 > /dev/null 2>&1
timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j2_stream.json 2> gpurun_out/j2_stream.err
grep -E "==|i32wrap" gpurun_out/j2_knn_variants.log | cut -c1-220
cat gpurun_out/j2_sort_ab.log; tail -6 gpurun_out/j2_build80.log
timeout 1500 python -u -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=30 > gpurun_out/j2_pytest.log 2>&1
tail -45 gpurun_out/j2_pytest.log
