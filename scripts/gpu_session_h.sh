# 1-GPU session H: host environment facts, full GPU tests, bench, BA launch list, C4
set -x
mkdir -p gpurun_out
(cat /sys/fs/cgroup/cpu.max; cat /sys/fs/cgroup/cpu.stat | head -6; python -c "import os;print('affinity',len(os.sched_getaffinity(0)))"; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)|MHz") > gpurun_out/hostinfo.txt 2>&1
cat gpurun_out/hostinfo.txt
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ba.csv python tests/gpu_ba_profile.py > gpurun_out/ba_prof.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/b_ncu4.log 2>&1
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
cat /sys/fs/cgroup/cpu.stat | head -6
