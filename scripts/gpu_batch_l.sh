set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cascade.py -q -m gpu -x > gpurun_out/pytest_cascade.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_cascade.log; tail -30 gpurun_out/pytest_cascade.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_cascade.py -q -m gpu -x -k "ragged or hash_tables" > gpurun_out/sanitizer_memcheck_cascade.log 2>&1; tail -5 gpurun_out/sanitizer_memcheck_cascade.log
