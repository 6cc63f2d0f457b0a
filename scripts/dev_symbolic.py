"""dev: phase timers of the symbolic analysis (ILUB_DEBUG=1) on 128^3 IChol0, warm allocator"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", 0)
A = P.poisson3d(128)
B = ilupp.IChol0Preconditioner(A); del B
torch.cuda.synchronize()
os.environ["ILUB_DEBUG"] = "1"
B = ilupp.IChol0Preconditioner(A)
print("timings_ms", B.timings_ms)
