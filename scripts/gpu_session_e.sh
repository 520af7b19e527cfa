# 1-GPU session E: full GPU tests, C2 bench, BA launch list + per-point Schur check, C3
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
R3D_BA_SCHUR=point timeout 600 python -m pytest tests/test_gpu_ba.py -x -q -m gpu > gpurun_out/pytest_ba_point.log 2>&1; tail -2 gpurun_out/pytest_ba_point.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ba.csv python tests/gpu_ba_profile.py > gpurun_out/ba_prof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu4.log 2>&1
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
