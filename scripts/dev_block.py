"""dev: block-sequential solve (ilub_block.cu) against the level loop (ilub_deep.cu) and the oracle on the irregular matrix.
usage: dev_block.py [n]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if os.environ.get("ILUB_PKG_ROOT"):
    sys.path.insert(0, os.environ["ILUB_PKG_ROOT"])
import torch
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
A = P.irregular_powerlaw(n=n)
b = P.rhs(n, 0)
L, U = oracle.port.ilu0(A)
ref = oracle.port.apply_lu(L, U, b)
reft = oracle.port.apply_lu(L, U, b, transpose=True)
ilupp.set_option("exact", 0)
for blk in (0, 1):
    B = ilupp.ILU0Preconditioner(A, options={"deep_block": blk})
    x0 = torch.from_numpy(b).cuda(); x = x0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(5):
        x.copy_(x0); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    y = x.cpu().numpy()
    z = b.copy(); B.apply_trans(z)
    print("deep_block %d: mode %s levels %s | apply %.2f ms | relerr apply %.2e apply_trans %.2e | symbolic %.0f ms" % (
        blk, B.stats["part_mode"], B.stats["levels"][:2], np.median(ts), np.abs(y - ref).max() / np.abs(ref).max(),
        np.abs(z - reft).max() / np.abs(reft).max(), B.timings_ms[0]), flush=True)
    if blk and os.environ.get("BLK_TRACE"):
        B.set_option("trace", 1); x.copy_(x0); B.apply(x); torch.cuda.synchronize(); B.set_option("trace", 0)
    del B
