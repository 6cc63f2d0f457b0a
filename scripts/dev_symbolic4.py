"""dev: phase timers (ILUB_DEBUG=1) of the creation of BASELINE config 4 (irregular power-law matrix, n = 10^6)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", int(sys.argv[1]) if len(sys.argv) > 1 else 0)
A = P.irregular_powerlaw()
B = ilupp.ILU0Preconditioner(A); del B
torch.cuda.synchronize()
os.environ["ILUB_DEBUG"] = "1"
B = ilupp.ILU0Preconditioner(A)
print("timings_ms", B.timings_ms)
