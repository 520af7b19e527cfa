# One GPU-box session: GPU parity tests, bench line, ncu captures (run: gpurun -- bash scripts/gpu_session.sh)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2_candidates_2sm -s 13 -c 1 -f -o gpurun_out/prof_k1_2sm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_f7_score -s 30 -c 1 -f -o gpurun_out/prof_f7_score python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bin_rerank -s 13 -c 1 -f -o gpurun_out/prof_bin_rerank python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/b_ncu4.log 2>&1
head -c 600 gpurun_out/bench.json
