for pdl in 0 1; do
  echo "== sys_pdl $pdl"
  python - <<PY
import sys, numpy as np, torch
sys.path.insert(0, '.')
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", 0); ilupp.set_option("sys_pdl", $pdl)
A = P.poisson3d(128); n = A.shape[0]
B = ilupp.IChol0Preconditioner(A)
x0 = torch.from_numpy(P.rhs(n, 0)).cuda(); x = x0.clone()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ts = []
for i in range(25):
    x.copy_(x0); torch.cuda.synchronize(); e0.record(); B.apply(x); e1.record(); torch.cuda.synchronize()
    if i >= 5: ts.append(e0.elapsed_time(e1))
print("single apply: median %.1f us min %.1f us" % (np.median(ts) * 1e3, min(ts) * 1e3))
# back-to-back applies (Krylov-like): 10 applies in one timed region
x.copy_(x0); torch.cuda.synchronize(); e0.record()
for _ in range(10): B.apply(x)
e1.record(); torch.cuda.synchronize()
print("10 applies back to back: %.1f us each" % (e0.elapsed_time(e1) * 100))
import oracle
Lo = oracle.port.ichol0(A); b = P.rhs(n, 0); yo = oracle.port.apply_llt(Lo, b); y = b.copy(); B.apply(y)
print("relerr", np.abs(y - yo).max() / np.abs(yo).max())
PY
done
