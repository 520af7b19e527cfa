set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -8 gpurun_out/pytest_gpu.log
for epi in 8 16; do
  R3D_K1_EPI=$epi timeout 300 python bench.py --workload c2-msurf64 --steps 5 --warmup 3 --no-ba --no-extras --no-filter --no-cpu-baseline > gpurun_out/bench_d64_epi$epi.json 2> gpurun_out/bench_d64_epi$epi.err
  python -c "import json; d=json.load(open('gpurun_out/bench_d64_epi$epi.json')); print('epi $epi', d['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'])"
done
R3D_DEBUG_TIMING=1 timeout 500 python bench.py --workload c2 --steps 3 --no-ba --no-extras > gpurun_out/bench_c2f.json 2> gpurun_out/bench_c2f.err; grep "r3d\] f" gpurun_out/bench_c2f.err | tail -4
python -c "import json; d=json.load(open('gpurun_out/bench_c2f.json')); print(json.dumps(d['f_filter']))"
R3D_DEBUG_TIMING=1 timeout 800 python bench.py --steps 2 --warmup 1 --no-ba --no-extras --no-cpu-baseline > gpurun_out/bench_c3f.json 2> gpurun_out/bench_c3f.err; grep "r3d\] f" gpurun_out/bench_c3f.err | tail -4
python -c "import json; d=json.load(open('gpurun_out/bench_c3f.json')); print(json.dumps(d['f_filter']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_acransac_fused -s 1 -c 1 -f -o gpurun_out/prof_acransac_fused4 python bench.py --workload c2 --steps 1 --warmup 1 --no-cpu-baseline --no-ba --no-extras > gpurun_out/b_ncu_f.log 2>&1; echo "ncu rc=$?"
