set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -k "test_gpu_ba or adaptor" > gpurun_out/pytest_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_j.log; tail -6 gpurun_out/pytest_j.log
for mode in env8 env1 dense; do
case $mode in env8) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=8;; env1) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=1;; dense) export R3D_BA_CHOL=dense;; esac
R3D_DEBUG_TIMING=1 timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/bench_ba_$mode.json 2> gpurun_out/bench_ba_$mode.err; grep "BA linear" gpurun_out/bench_ba_$mode.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/bench_ba_$mode.json'))['ba']; print('$mode', d['iters_per_s'], d['final_cost'], d['seconds_linear'])"
done
export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chol_envelope -s 3 -c 1 -f -o gpurun_out/prof_chol_env python bench.py --workload c2 --steps 1 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/b_ncu_env.log 2>&1; echo "ncu rc=$?"
unset R3D_BA_CHOL R3D_BA_ENV_CTAS
R3D_K1_DRAIN=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2_candidates_2sm -s 2 -c 1 -f -o gpurun_out/prof_k1_d64 python bench.py --workload c2-msurf64 --steps 1 --warmup 3 --no-ba --no-filter --no-extras --no-cpu-baseline > gpurun_out/b_ncu_d64.log 2>&1; echo "ncu rc=$?"
