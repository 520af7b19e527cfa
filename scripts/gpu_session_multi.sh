# multi-GPU box session: partitioned BA parity on 2 ranks, bench at N = 1, 2, 4 (weak scaling of the pair path,
# strong scaling of the BA leg)
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8; nproc
timeout 600 python -m pytest tests/test_gpu_ba_multi.py -q -m gpu > gpurun_out/pytest_ba_multi.log 2>&1; tail -3 gpurun_out/pytest_ba_multi.log
timeout 400 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 rc=$?"
for N in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err; echo "n$N rc=$?"
done
python - <<'PY'
import json
for n in (1,2,4):
    try:
        d=json.load(open("gpurun_out/scale_n%d.json"%n))
        print(n, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "ba", round(d["ba"]["iters_per_s"],1), d["ba"].get("n_gpus"), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "ERR", e)
PY
