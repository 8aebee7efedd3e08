"""Host-side multi-process logic on CPU: world_size=2, gloo backend, rendezvous on 127.0.0.1."""
import json
import os
import socket
import subprocess
import sys

import torch
import torch.multiprocessing as mp

from cmflow_b200 import dist as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 5, 256, 2048, 2049):
        for w in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


class _FakeNet:
    """Stands in for the CUDA model: deterministic per-pair outputs from the inputs, so sharded == unsharded is checkable on CPU."""
    def __call__(self, pc1, pc2, ft1, ft2, label, mode):
        B, _, N = pc1.shape
        sf = pc2 - pc1
        cls = torch.sigmoid(ft1[:, :1])
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, :3, 3] = sf.mean(-1)
        return sf, cls, T, cls.squeeze(1) > 0.5


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    w, r, _ = D.init(backend="gloo")
    assert (w, r) == (world, rank)
    g = torch.Generator().manual_seed(0)
    B = 5                                                    # odd: shards of 3 and 2
    pc1, pc2, ft1, ft2 = (torch.randn(B, 3, 16, generator=g) for _ in range(4))
    full = _FakeNet()(pc1, pc2, ft1, ft2, None, "test")
    got = D.sharded_forward(_FakeNet(), pc1, pc2, ft1, ft2, gather=True)
    ok = all(torch.equal(a, b) for a, b in zip(got, full))
    mx = D.reduce_max(10.0 + rank)                           # max over ranks of a per-rank elapsed time
    sm = D.reduce_sum(D.shard_bounds(B, world, rank)[1] - D.shard_bounds(B, world, rank)[0])
    # sharded evaluation: per-rank (batch-size weighted metric sums, pairs) -> one all-reduce -> the whole run's means (main_util.py:197-202)
    from cmflow_b200.eval_util import EvalAccumulator
    acc = EvalAccumulator({"r_res": 0.2, "theta_res": 0.026, "phi_res": 0.026}, "cpu")
    npairs = 3 if rank == 0 else 2
    acc.acc[:-1] = npairs * (1.0 + rank)                     # every metric = 1 on rank 0's 3 pairs, 2 on rank 1's 2 pairs
    acc.acc[-1] = npairs
    sf, seg, pose, n = acc.result()
    ev_ok = n == 5 and all(abs(v - 1.4) < 1e-12 for d in (sf, seg, pose) for v in d.values())
    if rank == 0:
        json.dump({"ok": ok and ev_ok, "max": mx, "sum": sm}, open(out, "w"))
    torch.distributed.destroy_process_group()


def test_two_rank_shard_gather_and_reductions(tmp_path):
    out = str(tmp_path / "r.json")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = json.load(open(out))
    assert r == {"ok": True, "max": 11.0, "sum": 5.0}


def test_bench_reference_arm_under_torchrun_prints_one_line_from_rank0():
    port = _free_port()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--points", "64"]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["value"] > 0
