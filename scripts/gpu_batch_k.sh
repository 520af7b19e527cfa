set -x
mkdir -p gpurun_out
bash scripts/gpu_session.sh smoke tests
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_default.err; head -c 600 gpurun_out/bench_default.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_n1.json | head -c 700; echo
timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-ba --no-extras > gpurun_out/c4_n1.json 2> gpurun_out/c4_n1.err; echo "c4 rc=$?"; head -c 400 gpurun_out/c4_n1.json; echo
bash scripts/gpu_session.sh ncu_k1 ncu_launches
