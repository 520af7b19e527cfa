set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -k "test_gpu_ba or test_gpu_match or golden or adaptor" > gpurun_out/pytest_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_h.log; tail -6 gpurun_out/pytest_h.log
for mode in env8 env4 env1 dense; do
case $mode in env8) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=8;; env4) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=4;; env1) export R3D_BA_CHOL=envelope R3D_BA_ENV_CTAS=1;; dense) export R3D_BA_CHOL=dense;; esac
R3D_DEBUG_TIMING=1 timeout 600 python bench.py --workload c2 --steps 2 --warmup 3 --no-filter --no-extras --no-cpu-baseline > gpurun_out/bench_ba_$mode.json 2> gpurun_out/bench_ba_$mode.err; grep "BA linear" gpurun_out/bench_ba_$mode.err | tail -1
python -c "import json; d=json.load(open('gpurun_out/bench_ba_$mode.json'))['ba']; print('$mode', d['iters_per_s'], d['final_cost'], d['seconds_linear'])"
done
unset R3D_BA_CHOL R3D_BA_ENV_CTAS
for cfg in "0 8" "1 8" "1 16" "0 16"; do
set -- $cfg
for wl in c2-msurf64 c2 c3; do
R3D_K1_DRAIN=$1 R3D_K1_EPI=$2 timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-ba --no-extras --no-filter --no-cpu-baseline > gpurun_out/bench_dr$1_e$2_$wl.json 2> gpurun_out/bench_dr$1_e$2_$wl.err
python -c "import json; d=json.load(open('gpurun_out/bench_dr$1_e$2_$wl.json')); print('drain $1 epi $2 $wl', round(d['value']), round(d['roofline']['frac'],3), round(d['breakdown_ms']['candidates'],2), d['clocks']['sm_mhz'])"
done
done
