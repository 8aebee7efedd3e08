"""Developer probe (not a test): per-role barrier wait cycles of the fused set-conv #2 kernel at the bench point (CMF_SC2_DBG=1; with a
library built with -DSC2_PROD_PROFILE also the sections of a producer iteration).   python tests/sc2_wait_probe.py"""
import os, sys, torch
sys.path.insert(0, ".")
from cmflow_b200.cmflow import CMFlow
from cmflow_b200.synth import make_pairs, synthetic_state_dict
class A: num_points=256; stat_thres=0.5
net=CMFlow(A()); net.load_state_dict(synthetic_state_dict(0)); net=net.cuda(); net.set_precision("fp16x3")
inp=[t.cuda() for t in make_pairs(256,256,seed=1)[:4]]
for i in range(3):
    with torch.no_grad(): net(*inp,None,"test")
torch.cuda.synchronize()
os.environ["CMF_SC2_DBG"]="1"
with torch.no_grad(): net(*inp,None,"test")
torch.cuda.synchronize()
