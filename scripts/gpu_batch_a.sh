set -x
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
R3D_DEBUG_TIMING=1 timeout 500 python bench.py --workload c2 --steps 3 --no-ba --no-extras > gpurun_out/bench_c2f.json 2> gpurun_out/bench_c2f.err; grep "r3d\] f" gpurun_out/bench_c2f.err | tail -12
python -c "import json; d=json.load(open('gpurun_out/bench_c2f.json')); print(json.dumps(d['f_filter']))"
R3D_DEBUG_TIMING=1 timeout 800 python bench.py --steps 2 --warmup 1 --no-ba --no-extras --no-cpu-baseline > gpurun_out/bench_c3f.json 2> gpurun_out/bench_c3f.err; grep "r3d\] f" gpurun_out/bench_c3f.err | tail -12
python -c "import json; d=json.load(open('gpurun_out/bench_c3f.json')); print(json.dumps(d['f_filter']))"
