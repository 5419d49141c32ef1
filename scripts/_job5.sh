set -x
mkdir -p gpurun_out; rm -f gpurun_out/j5_*
timeout 900 python -u -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 200 --durations=8 > gpurun_out/j5_pytest.log 2>&1
tail -14 gpurun_out/j5_pytest.log
timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j5_stream.json 2> gpurun_out/j5_stream.err
python -c "
import json;d=json.load(open('gpurun_out/j5_stream.json'))
for k,v in d['stages'].items(): print(k, v['ms'], v['frac_of_hbm_peak'])"
GSX_LIB=$PWD/3dgsconverter_b200/lib/variants/libgsx_onepass.so timeout 200 python scripts/stream_kernels_probe.py 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read())
print('ONEPASS compact', d['stages']['compact_points(50%)'])"
GSX_LIB=$PWD/3dgsconverter_b200/lib/variants/libgsx_onepass.so timeout 200 python -u -m pytest tests/test_dropin_api_gpu.py -x -q -m gpu -p no:cacheprovider --timeout 100 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/j5_bench_n1.json 2> gpurun_out/j5_bench_n1.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/j5_launches.log 2>&1
timeout 400 python bench.py --config c3 > gpurun_out/j5_bench_c3.json 2> gpurun_out/j5_bench_c3.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/j5_bench_ref.json 2> gpurun_out/j5_bench_ref.err
cut -c1-400 gpurun_out/j5_bench_n1.json; cut -c1-300 gpurun_out/j5_bench_c3.json; cut -c1-300 gpurun_out/j5_bench_ref.json
