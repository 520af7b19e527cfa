# final 1-GPU session of the round: smoke, full GPU tests, bench (both arms), ncu evidence, C3 / C4 for the record
set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2_candidates_2sm -s 13 -c 1 -f -o gpurun_out/prof_k1_2sm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bin_rerank -s 13 -c 1 -f -o gpurun_out/prof_bin_rerank python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/b_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_ba.csv python tests/gpu_ba_profile.py > gpurun_out/ba_prof.log 2>&1
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 rc=$?"
timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-ba --no-filter --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?"
