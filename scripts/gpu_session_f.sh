# 1-GPU session F: full GPU tests with the new K1 options, A/B of the candidate-kernel switches, bench, ncu of K1
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for cfg in "1 2" "0 2" "1 1" "0 1"; do
  set -- $cfg
  R3D_K1_VOTE=$1 R3D_K1_QBUF=$2 timeout 300 python bench.py --steps 6 --no-cpu-baseline --no-ba --no-filter > gpurun_out/bench_v$1_q$2.json 2> gpurun_out/bench_v$1_q$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_v$1_q$2.json"))
print("vote=$1 qbuf=$2", round(d["value"]), round(d["ms_per_step"],1), {k:round(v,2) for k,v in d["breakdown_ms"].items()}, round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
PY
done
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_l2_candidates_2sm -s 13 -c 1 -f -o gpurun_out/prof_k1_2sm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ba > gpurun_out/b_ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bin_rerank -s 13 -c 1 -f -o gpurun_out/prof_bin_rerank python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter > gpurun_out/b_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ba_schur_batched -s 2 -c 1 -f -o gpurun_out/prof_schur python tests/gpu_ba_profile.py > gpurun_out/b_ncu5.log 2>&1
tail -2 gpurun_out/b_ncu5.log
