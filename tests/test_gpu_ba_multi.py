"""-m gpu: multi-rank bundle adjustment.  One process per GPU under torch.distributed.run;
the reduced camera systems of the point partitions are summed by the library's own ncclAllReduce.  World 1
exercises the NCCL path on a single-GPU box; world 2 needs two GPUs (skipped otherwise)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [1, 2])
def test_partitioned_ba_equals_single_gpu_and_oracle(world, tmp_path, r3dlib, oracle):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    out = tmp_path / "r.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "mgpu_ba_worker.py"), str(out), "12", "3000", "15"]
    env = dict(os.environ, NCCL_DEBUG="WARN")
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    r = json.load(open(out))
    assert r["world"] == world and r["cams_identical"]
    assert r["iterations"][0] == r["iterations"][1] == r["iterations"][2]
    assert r["successful"][0] == r["successful"][1] == r["successful"][2]
    assert r["trace_vs_single"] < 1e-8 and r["trace_vs_oracle"] < 1e-8
    assert r["residual_rel_vs_oracle"] < 1e-5          # north_star bar for BA
    assert r["final_cost"] < 0.5 * r["initial_cost"]          # 2 % gross outliers keep a Huber floor
