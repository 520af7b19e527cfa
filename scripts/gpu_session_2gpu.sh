# 2-GPU box session: BA parity (single + partitioned), 1- and 2-GPU bench lines
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ba.py tests/test_gpu_ba_multi.py -x -q -m gpu > gpurun_out/pytest_ba.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ba.log
tail -15 gpurun_out/pytest_ba.log
timeout 600 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/bench_g1.json 2> gpurun_out/bench_g1.err; echo "bench1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "bench2 rc=$?"
tail -3 gpurun_out/bench_g2.err
