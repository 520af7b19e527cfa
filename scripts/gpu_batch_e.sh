# usage: bash scripts/gpu_batch_e.sh N   (run under gpurun --gpus N)
set -x
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -q -m gpu -k "multidevice or ba_multi or multi" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log; tail -6 gpurun_out/pytest_multi.log
run() {  # run <n> <tag> <args...>
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" > gpurun_out/scale_${tag}_n$n.json 2> gpurun_out/scale_${tag}_n$n.err; fi
  echo "rc=$? n=$n $tag"; tail -2 gpurun_out/scale_${tag}_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_${tag}_n$n.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('impl','value','ms_per_step','n_gpus')}, 'e2e', d.get('e2e',{}).get('value'), 'filter', (d.get('f_filter') or {}).get('pairs_per_s'), 'ba', (d.get('ba') or {}).get('value'))
except Exception as e: print('parse failed', e)
PY
}
for n in 1 $N; do
  run $n own --steps 3 --warmup 3
done
run $N ref --impl reference --steps 1 --warmup 0
