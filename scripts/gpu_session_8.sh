# 8-GPU box: the scaling end point (weak scaling of the pair path, strong scaling of the BA leg)
set -x
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max; nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/scale_n8.json 2> gpurun_out/scale_n8.err; echo "n8 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/scale_n8.json"))
print(8, round(d["value"]), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]), "ba", round(d["ba"]["iters_per_s"],1), d["ba"].get("n_gpus"), "clk", d["clocks"], {k:round(v,1) for k,v in d["breakdown_ms"].items()})
PY
tail -3 gpurun_out/scale_n8.err
