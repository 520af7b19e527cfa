set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -k "ba or adaptor or sfm or structure" > gpurun_out/pytest_ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ba.log; tail -8 gpurun_out/pytest_ba.log
for mode in env8 env1 dense; do
case $mode in env8) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=8;; env1) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=1;; dense) export R3D_BA_CHOL=dense;; esac
R3D_DEBUG_TIMING=1 timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/bench_ba_$mode.json 2> gpurun_out/bench_ba_$mode.err; grep "BA linear" gpurun_out/bench_ba_$mode.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/bench_ba_$mode.json'))['ba']; print('$mode', d['iters_per_s'], d['final_cost'], d['seconds_linear'])"
done
unset R3D_BA_CHOL R3D_BA_ENV_CTAS
R3D_LIB=$PWD/regard3d_b200/libr3dgpu_c16.so timeout 900 python -m pytest tests -q -m gpu -x -k "test_gpu_match or golden" > gpurun_out/pytest_c16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c16.log; tail -4 gpurun_out/pytest_c16.log
for lib in libr3dgpu.so libr3dgpu_c16.so; do
for wl in c2-msurf64 c2 c3; do
R3D_LIB=$PWD/regard3d_b200/$lib timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-ba --no-extras --no-filter --no-cpu-baseline > gpurun_out/bench_ab_${lib}_$wl.json 2> gpurun_out/bench_ab_${lib}_$wl.err
python -c "import json; d=json.load(open('gpurun_out/bench_ab_${lib}_$wl.json')); print('$lib $wl', round(d['value']), round(d['roofline']['frac'],3), d['breakdown_ms'], d['result'].get('stage_b_query_frac'), d['result'].get('early_rejected_query_frac'))"
done
done
