# One GPU-box session (run: gpurun --timeout T -- bash scripts/gpu_session.sh STEP [STEP...]); outputs in gpurun_out/.
# Steps: smoke tests bench bench_ref sanitize ncu_k1 ncu_launches ncu_ba c2 c4
set -x
mkdir -p gpurun_out
for step in "$@"; do
case $step in
smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log ;;
tests) timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
bench) timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json ;;
bench_ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json ;;
c2) timeout 600 python bench.py --workload c2 --no-ba > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?" ;;
c4) timeout 900 python bench.py --workload c4 --steps 1 --warmup 1 --no-ba > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?" ;;
sanitize)
  for tool in memcheck racecheck synccheck; do
    for part in match filter ba liop; do
      timeout 200 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py $part > gpurun_out/sanitizer_${tool}_${part}.log 2>&1
      echo "$tool $part rc=$?"; tail -3 gpurun_out/sanitizer_${tool}_${part}.log
    done
  done ;;
ncu_k1) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_l2_candidates_2sm -s 4 -c 1 -f -o gpurun_out/prof_k1_2sm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-filter --no-extras ${NCU_BENCH_ARGS:-} > gpurun_out/b_ncu.log 2>&1; echo "ncu_k1 rc=$?" ;;
ncu_launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-extras ${NCU_BENCH_ARGS:-} > gpurun_out/b_ncu4.log 2>&1; echo "ncu_launches rc=$?" ;;
ncu_ba) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ba.csv python tests/gpu_ba_profile.py > gpurun_out/ba_prof.log 2>&1; echo "ncu_ba rc=$?" ;;
filter) timeout 600 python -m pytest tests -x -q -m gpu -k "filter or compute_matches or golden" > gpurun_out/pytest_filter.log 2>&1; tail -15 gpurun_out/pytest_filter.log
  R3D_DEBUG_TIMING=1 timeout 500 python bench.py --workload c2 --steps 3 --no-ba --no-extras > gpurun_out/bench_c2f.json 2> gpurun_out/bench_c2f.err; tail -5 gpurun_out/bench_c2f.err
  python -c "import json; d=json.load(open('gpurun_out/bench_c2f.json')); print(json.dumps(d['f_filter']))" ;;
ncu_filter) timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_acransac_fused -s 1 -c 1 -f -o gpurun_out/prof_acransac_fused python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-extras > gpurun_out/b_ncu_f.log 2>&1; echo "ncu_filter rc=$?" ;;
*) echo "unknown step $step" ;;
esac
done
