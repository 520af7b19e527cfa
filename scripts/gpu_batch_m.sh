set -x
mkdir -p gpurun_out
R3D_DEBUG_TIMING=1 timeout 900 python bench.py --workload c4 --steps 2 --warmup 3 --no-ba --no-extras > gpurun_out/c4c_n1.json 2> gpurun_out/c4c_n1.err; echo "c4 rc=$?"; tail -5 gpurun_out/c4c_n1.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c4c_n1.json'))
print(d['value'], d['e2e']['value'], d['ms_per_step']); print(d['roofline']); print(d['breakdown_ms']); print(d['result']); print(d.get('cpu_baseline')); print((d.get('f_filter') or {}).get('pairs_per_s'))
PY
timeout 600 python bench.py --impl reference --workload c4 --steps 1 --warmup 0 > gpurun_out/c4c_ref_n1.json 2> gpurun_out/c4c_ref_n1.err; echo "ref rc=$?"; head -c 900 gpurun_out/c4c_ref_n1.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_cascade_match -s 3 -c 1 -f -o gpurun_out/prof_cascade_match python bench.py --workload c4 --steps 1 --warmup 3 --no-ba --no-extras --no-filter --no-cpu-baseline > gpurun_out/b_ncu_cascade.log 2>&1; echo "ncu rc=$?"
