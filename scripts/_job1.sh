set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/j1_pytest.log
timeout 200 python scripts/build_profile.py 80000000 > gpurun_out/j1_build80.log 2>&1
timeout 200 python scripts/stream_kernels_probe.py > gpurun_out/j1_stream.json 2> gpurun_out/j1_stream.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_sor_knn -c 1 -f -o gpurun_out/r02_knn python bench.py --steps 1 --warmup 1 --no-extras > gpurun_out/j1_ncu_knn.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_steps2.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/j1_launches.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck_smoke.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r02_sanitizer_memcheck_smoke.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck_smoke.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r02_sanitizer_racecheck_smoke.log
timeout 300 python bench.py > gpurun_out/j1_bench_n1.json 2> gpurun_out/j1_bench_n1.err
tail -3 gpurun_out/j1_pytest.log; cat gpurun_out/j1_build80.log | tail -6; tail -2 gpurun_out/r02_sanitizer_*.log
