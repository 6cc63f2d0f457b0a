"""dev: phase timers (ILUB_DEBUG=1) of the creation of the batched config (4096 x 128^2 ILU0), warm allocator"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
nsys = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
ilupp.set_option("exact", 0)
pat, vals = P.batched_poisson2d(nsys, 128)
B = ilupp.ILU0PreconditionerBatched(pat, vals); print("first timings_ms", B.timings_ms); del B
torch.cuda.synchronize()
os.environ["ILUB_DEBUG"] = "1"
B = ilupp.ILU0PreconditionerBatched(pat, vals)
print("timings_ms", B.timings_ms)
