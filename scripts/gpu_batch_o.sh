# usage: bash scripts/gpu_batch_o.sh   (run under gpurun --gpus 8)
set -x
mkdir -p gpurun_out
tr() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n "$@"; }
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ('impl','value','ms_per_step','n_gpus')}, 'e2e', d.get('e2e',{}).get('value'), 'filter', (d.get('f_filter') or {}).get('pairs_per_s'), 'gather', (d.get('result') or {}).get('gather_ms'), 'bd', d.get('breakdown_ms'), 'roof', (d.get('roofline') or {}).get('frac'))
except Exception as e: print('parse failed', sys.argv[1], e)
PY
}
tr 8 --workload c4 --steps 2 --warmup 3 --no-ba --no-extras > gpurun_out/c4c_n8.json 2> gpurun_out/c4c_n8.err; echo rc=$?; grep -i "gather\|error" gpurun_out/c4c_n8.err | tail -3; show gpurun_out/c4c_n8.json
tr 8 --workload c4 --impl reference --steps 1 --warmup 0 > gpurun_out/c4c_ref_n8.json 2> gpurun_out/c4c_ref_n8.err; echo rc=$?; show gpurun_out/c4c_ref_n8.json
tr 4 --steps 3 --warmup 3 --no-ba --no-extras --no-cpu-baseline > gpurun_out/scale4_own.json 2> gpurun_out/scale4_own.err; echo rc=$?; grep -i "gather\|error" gpurun_out/scale4_own.err | tail -3; show gpurun_out/scale4_own.json
tr 2 --steps 3 --warmup 3 --no-ba --no-extras --no-cpu-baseline > gpurun_out/scale2_own.json 2> gpurun_out/scale2_own.err; echo rc=$?; show gpurun_out/scale2_own.json
tr 8 --steps 3 --warmup 3 --no-extras > gpurun_out/scale8_own.json 2> gpurun_out/scale8_own.err; echo rc=$?; show gpurun_out/scale8_own.json
