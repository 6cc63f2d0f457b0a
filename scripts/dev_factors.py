"""dev: GPU apply of reference ILUT / ICholT factors under the different solve schedules.  usage: dev_factors.py N"""
import os, sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ilupp_b200 as ilupp
import oracle
from ilupp_b200 import problems as P

N = int(sys.argv[1])
m = oracle.ref_module()
A = P.poisson3d(N); A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
n = A.shape[0]
ilupp.set_option("exact", int(os.environ.get("ILUB_EXACT", "0")))
for name, make in (("ILUT", lambda: m.ILUTPreconditioner(A.data, A.indices, A.indptr, True, 10, 1e-3)),
                   ("ICholT", lambda: m.ICholTPreconditioner(A.data, A.indices, A.indptr, True, 5, 1e-3))):
    R = make()
    fs = [(sp.csr_matrix if f[3] else sp.csc_matrix)((f[0], f[1], f[2]), shape=(n, n)) for f in R.factors_info()]
    b = P.rhs(n, 0)
    yr = b.copy(); R.apply(yr)
    for algo in ((0,) if os.environ.get('ILUB_EXACT') else (0, 1, 2, 3)):
        ilupp.set_option("solve_algo", algo)
        try:
            B = ilupp.FactorsPreconditioner(fs[0], fs[1] if len(fs) > 1 else None)
            x0 = torch.from_numpy(b).cuda(); x = x0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for i in range(5):
                x.copy_(x0); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
                if i >= 2: ts.append(e0.elapsed_time(e1))
            err = np.abs(x.cpu().numpy() - yr).max() / np.abs(yr).max()
            print("%s N=%d solve_algo=%d: apply %.2f ms relerr %.1e levels %s mode %s parts %s" % (name, N, algo, np.median(ts), err, B.stats["levels"][:2], B.stats["part_mode"], B.stats["nparts"]), flush=True)
            del B
        except Exception as e:
            print(name, algo, "failed:", e, flush=True)
    ilupp.set_option("solve_algo", 0)
