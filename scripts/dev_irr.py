"""dev: where does the factorisation time of the irregular power-law matrix go?  numeric time against the row-length cap"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", 0)
for max_len in (32, 128, 512, 4096):
    A = P.irregular_powerlaw(max_len=max_len)
    lens = np.diff(A.indptr)
    B = ilupp.ILU0Preconditioner(A)
    st = B.stats
    x = torch.from_numpy(P.rhs(A.shape[0], 0)).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    B.apply(x); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
    print("max_len %4d: nnz %.1fM, rows > 64: %d, > 512: %d, longest %d | levels %s | symbolic %.0f numeric %.0f ms | apply %.1f ms" % (
        max_len, A.nnz / 1e6, (lens > 64).sum(), (lens > 512).sum(), lens.max(), st["levels"][:2], B.timings_ms[0], B.timings_ms[1], e0.elapsed_time(e1)), flush=True)
    del B
