"""`import pointnet2_cuda` lands here when cmflow_b200/shim is on sys.path."""
from cmflow_b200.pointnet2_cuda import *  # noqa: F401,F403
from cmflow_b200.pointnet2_cuda import __all__  # noqa: F401
