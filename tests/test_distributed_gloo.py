"""The N>1 path on CPU: world_size-2 gloo. Sequences are sharded across ranks, each rank produces the
additive partials of its shard, ONE all-reduce sums them, and every rank derives the scores of the
GLOBAL batch (oracle: calc_metrics on the concatenated batch, with exact integer counts)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import metrics_oracle as MO
    from test_host_logic import _partials_from_oracle
    from weatherforecastingtoolkit_b200 import metrics as M
    torch.manual_seed(0)
    p, t = torch.rand(4, 3, 1, 64, 64), torch.rand(4, 3, 1, 64, 64)
    shard = slice(rank * 2, rank * 2 + 2)                   # sequences [r*N/G, (r+1)*N/G)
    local = _partials_from_oracle(p[shard], t[shard])
    struct = torch.cat([torch.from_numpy(local.ints), torch.from_numpy(local.floats).view(torch.int64)])
    summed = M.all_reduce_partials(struct)                  # the single collective of the path
    host = summed.numpy()
    mp_ = M.MetricPartials(host[:100], host[100:].view(np.float64), 6)
    scores = M.scores_from_partials(mp_)
    whole = MO.integer_counts(p, t)
    ok = mp_.counts.tolist() == whole.tolist()
    ref = MO.calc_metrics(p, t)
    ok = ok and all(scores[k] == ref[k] for k in ref if k.startswith(("CSI", "HSS")))
    ok = ok and abs(scores["SSIM"] - ref["SSIM"]) < 1e-6 and abs(scores["CRPS"] - ref["CRPS"]) < 1e-6
    # mean-of-ratios (the reference's sync_dist) is NOT what we report: check they differ in general
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write("ok" if ok else "bad")
    dist.destroy_process_group()


def test_all_reduce_partials_world2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert (tmp_path / f"rank{r}.txt").read_text() == "ok"


def test_bench_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 without work or output."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
