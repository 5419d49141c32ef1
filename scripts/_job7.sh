set -x
mkdir -p gpurun_out; rm -f gpurun_out/j7_*
timeout 900 python -u -m pytest tests -x -q -m gpu -p no:cacheprovider --timeout 200 --durations=5 > gpurun_out/j7_pytest.log 2>&1
tail -12 gpurun_out/j7_pytest.log
