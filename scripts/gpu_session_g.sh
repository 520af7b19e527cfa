# 1-GPU session G: full GPU tests, bench (+debug timing), launch lists
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
R3D_DEBUG_TIMING=1 timeout 600 python bench.py --steps 3 --no-cpu-baseline --no-ba > gpurun_out/bench_dbg.json 2> gpurun_out/bench_dbg.err
grep "r3d\]" gpurun_out/bench_dbg.err | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ba.csv python tests/gpu_ba_profile.py > gpurun_out/ba_prof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/b_ncu4.log 2>&1
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
