# One GPU-box session (run: gpurun --timeout T -- bash scripts/gpu_session.sh STEP [STEP...]); outputs in gpurun_out/.
# Steps: smoke tests bench bench_ref sanitize ncu_k1 ncu_launches ncu_ba c2 c4 c4exact filter ncu_filter ncu_cascade
#        ab_epilogue ab_chol scale (multi-GPU: gpurun --gpus N, SCALE_N="1 2 4 8")
set -x
mkdir -p gpurun_out
for step in "$@"; do
case $step in
smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log ;;
tests) timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
bench) timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json ;;
bench_ref) timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json ;;
c2) timeout 600 python bench.py --workload c2 --no-ba > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "c2 rc=$?" ;;
c4) timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-ba --no-extras > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "c4 rc=$?" ;;
c4exact) timeout 900 python bench.py --workload c4-exact --steps 2 --warmup 3 --no-ba --no-extras > gpurun_out/bench_c4_exact.json 2> gpurun_out/bench_c4_exact.err; echo "c4exact rc=$?" ;;
ncu_cascade) timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cascade_match -s 3 -c 1 -f -o gpurun_out/prof_cascade_match python bench.py --workload c4 --steps 1 --warmup 3 --no-ba --no-extras --no-filter --no-cpu-baseline > gpurun_out/b_ncu_cascade.log 2>&1; echo "ncu_cascade rc=$?" ;;
ab_epilogue)  # candidate-kernel epilogue A/B (profiles/r02_chunk_ab.md): TMEM drain order x epilogue warps, three workloads
  for cfg in "0 8" "1 8" "1 16" "0 16"; do set -- $cfg
    for wl in c2-msurf64 c2 c3; do
      R3D_K1_DRAIN=$1 R3D_K1_EPI=$2 timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-ba --no-extras --no-filter --no-cpu-baseline > gpurun_out/bench_dr$1_e$2_$wl.json 2> gpurun_out/bench_dr$1_e$2_$wl.err
      python -c "import json; d=json.load(open('gpurun_out/bench_dr$1_e$2_$wl.json')); print('drain $1 epi $2 $wl', round(d['value']), round(d['roofline']['frac'],3), round(d['breakdown_ms']['candidates'],2))"
    done
  done ;;
ab_chol)  # BA linear solve A/B (profiles/r02_ba_cholesky_ab.md)
  for mode in "envelope 8" "envelope 1" "dense 8"; do set -- $mode
    R3D_BA_CHOL=$1 R3D_BA_ENV_CTAS=$2 R3D_DEBUG_TIMING=1 timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/bench_ba_$1$2.json 2> gpurun_out/bench_ba_$1$2.err
    python -c "import json; d=json.load(open('gpurun_out/bench_ba_$1$2.json'))['ba']; print('$1 $2', d['iters_per_s'], d['final_cost'], d['seconds_linear'])"
  done ;;
scale)  # strong scaling of the default workload + the reference arm, one torchrun per N (gpurun --gpus max(N))
  for n in ${SCALE_N:-1 2 4 8}; do
    if [ "$n" = 1 ]; then run="python"; else run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"; fi
    timeout 900 $run bench.py --gpus $n ${BENCH_ARGS:---steps 3 --warmup 3} > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; echo "scale n=$n rc=$?"
    timeout 600 $run bench.py --gpus $n --impl reference --steps 1 --warmup 0 > gpurun_out/scale_ref_n$n.json 2> gpurun_out/scale_ref_n$n.err; echo "ref n=$n rc=$?"
    python -c "import json; d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['e2e']['value'], (d['result'] or {}).get('gather_ms'))"
  done ;;
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
