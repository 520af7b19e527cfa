# usage: bash scripts/gpu_batch_i.sh   (run under gpurun --gpus 8)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; cat /sys/fs/cgroup/cpu.max
tr() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $n "$@"; }
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ('impl','value','ms_per_step','n_gpus')}, 'e2e', d.get('e2e',{}).get('value'), 'filter', (d.get('f_filter') or {}).get('pairs_per_s'), 'ba', (d.get('ba') or {}).get('iters_per_s'), 'gather', (d.get('result') or {}).get('gather_ms'), 'bd', d.get('breakdown_ms'))
except Exception as e: print('parse failed', sys.argv[1], e)
PY
}
tr 8 --steps 3 --warmup 3 > gpurun_out/scale8_own.json 2> gpurun_out/scale8_own.err; echo rc=$?; tail -3 gpurun_out/scale8_own.err; show gpurun_out/scale8_own.json
tr 8 --impl reference --steps 1 --warmup 0 > gpurun_out/scale8_ref.json 2> gpurun_out/scale8_ref.err; echo rc=$?; show gpurun_out/scale8_ref.json
tr 8 --workload c4 --steps 2 --warmup 3 --no-ba --no-extras > gpurun_out/c4_n8.json 2> gpurun_out/c4_n8.err; echo rc=$?; tail -3 gpurun_out/c4_n8.err; show gpurun_out/c4_n8.json
tr 4 --steps 3 --warmup 3 --no-ba --no-extras --no-cpu-baseline > gpurun_out/scale4_own.json 2> gpurun_out/scale4_own.err; echo rc=$?; show gpurun_out/scale4_own.json
R3D_GATHER=p2p tr 8 --steps 3 --warmup 3 --no-ba --no-extras --no-cpu-baseline --no-filter > gpurun_out/scale8_p2p.json 2> gpurun_out/scale8_p2p.err; echo rc=$?; show gpurun_out/scale8_p2p.json
