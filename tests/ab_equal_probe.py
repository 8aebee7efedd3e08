"""Developer probe (not a test): do two builds of the library give the same bits?   python tests/ab_equal_probe.py out.pt   (run once per CMF_LIB, then compare)"""
import sys, torch
sys.path.insert(0, ".")
from cmflow_b200.cmflow import CMFlow
from cmflow_b200.synth import make_pairs, synthetic_state_dict
class A: num_points = 256; stat_thres = 0.5
net = CMFlow(A()); net.load_state_dict(synthetic_state_dict(0)); net = net.cuda(); net.set_precision("fp16x3")
res = {}
for B, N in ((40, 256), (3, 200)):
    inp = [t.cuda() for t in make_pairs(B, N, seed=7)[:4]]
    with torch.no_grad(): out = net(*inp, None, "test")
    torch.cuda.synchronize()
    res[(B, N)] = [o.cpu() for o in out] + [net.tap("cost1", (B * N, 512)).cpu(), net.tap("prop", (B * N, 256)).cpu()]
if len(sys.argv) > 2:
    ref = torch.load(sys.argv[2])
    for k in res:
        print(k, [bool(torch.equal(a, b)) for a, b in zip(res[k], ref[k])])
torch.save(res, sys.argv[1])
