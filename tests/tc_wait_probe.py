"""Developer probe (not a test): per-role barrier wait cycles of the CTA-pair tensor-core GEMM on a large plain GEMM.
   python tests/tc_wait_probe.py [M K cols [fmt]]      fmt 0 = 3xTF32, 1 = 3xFP16"""
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tests/", 1)[0])
from cmflow_b200._lib import check, dptr, lib, stream_ptr  # noqa: E402

M, K, cols = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 512, 2_000_000)
fmt = int(sys.argv[4]) if len(sys.argv) > 4 else 1
SK = 32 if fmt else 16
dev = "cuda"
W = torch.randn(M, K, device=dev) / K ** 0.5
X = torch.randn(cols, K, device=dev)
out = torch.empty(cols, M, device=dev)
scratch = torch.empty(lib().cmf_test_tc_tiled_floats(M, K), device=dev)
dbg = torch.zeros(148, 8, dtype=torch.int64, device=dev)
for it in range(3):
    lib().cmf_test_tc_set_dbg(dptr(dbg) if it == 2 else None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib().cmf_test_tc_gemm_fmt(fmt, M, K, cols, dptr(W), K, dptr(X), K, None, 1, dptr(out), M, dptr(scratch), 0, None, None, stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"fmt {fmt} M {M} K {K} cols {cols} run {it}: {ms:.3f} ms  {2*M*K*cols/ms/1e9:.1f} TFLOP/s algorithmic")
lib().cmf_test_tc_set_dbg(None)
d = dbg.cpu().double()
names = ["total", "mma:tempty", "mma:full", "mma:peer_full", "loader:empty", "producer:empty", "epilogue:tfull", "tiles"]
lead, peer = d[0::2], d[1::2]
for i, n in enumerate(names):
    print(f"{n:16s} leader mean {lead[:, i].mean():12.0f}  peer mean {peer[:, i].mean():12.0f}")
tiles = lead[:, 7].mean()
stages = tiles * (K // SK)
print(f"per stage: total {lead[:,0].mean()/stages:.0f} clk; mma waits: tempty {lead[:,1].mean()/stages:.0f} full {lead[:,2].mean()/stages:.0f} peer_full {lead[:,3].mean()/stages:.0f}; "
      f"loader empty {lead[:,4].mean()/stages:.0f}; producer empty {lead[:,5].mean()/stages:.0f}; epilogue tfull/tile {lead[:,6].mean()/tiles:.0f}")
