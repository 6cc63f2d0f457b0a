#!/bin/bash
# compute-sanitizer evidence: memcheck + racecheck over the kernel families on small problems
# (the sync-free kernels hand values between CTAs through global mailboxes on purpose: racecheck only sees shared memory,
# which is where the rings, windows and TMA stages live).  usage (under gpurun): scripts/sanitize.sh -> gpurun_out/sanitizer_*.log
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import ilupp_b200 as ilupp
from ilupp_b200 import problems as P
import oracle
case = sys.argv[1]
ilupp.set_option("spin_limit_ms", 60000)   # the tools slow kernels down by orders of magnitude
if case == "systolic":      # k_trsv_systolic (3-D, clusters + DSMEM + mailboxes), stencil factorisation on the same kernel
    ilupp.set_option("exact", 0); A = P.poisson3d(20); B = ilupp.IChol0Preconditioner(A)
elif case == "systolic_exact":
    ilupp.set_option("exact", 1); A = P.poisson3d(16); B = ilupp.IChol0Preconditioner(A)
elif case == "plane":       # k_trsv_plane: batch of 2-D systems, 4 right-hand sides
    ilupp.set_option("exact", 0); pat, vals = P.batched_poisson2d(6, 40); B = ilupp.ILU0PreconditionerBatched(pat, vals); A = None
elif case == "syncfree":    # k_trsv_syncfree + sync-free factorisation k_factor on a random matrix
    ilupp.set_option("exact", 1); ilupp.set_option("solve_algo", 2); ilupp.set_option("factor_algo", 2)
    import scipy.sparse as sp
    R = sp.random(3000, 3000, density=0.002, random_state=1, format="csr"); A = (R + R.T + 10 * sp.eye(3000)).tocsr(); A.sort_indices()
    A.indices = A.indices.astype(np.int32); A.indptr = A.indptr.astype(np.int32)
    B = ilupp.ILU0Preconditioner(A)
elif case == "levels":      # k_factor_levels + k_trsv_levels (cooperative level loops) and k_trsv_level_packed
    ilupp.set_option("exact", 0); ilupp.set_option("solve_algo", 4)
    A = P.poisson3d(14); perm = np.random.default_rng(1).permutation(A.shape[0]); A = A[perm][:, perm].tocsr(); A.sort_indices()
    B = ilupp.ILU0Preconditioner(A)
elif case == "deep":        # k_trsv_deep (one-CTA level loop) on a small irregular power-law matrix
    ilupp.set_option("exact", 0); A = P.irregular_powerlaw(n=6000); B = ilupp.ILU0Preconditioner(A)
n = B.shape[0]
b = P.rhs(n, 0)
if case == "plane":
    X = np.asfortranarray(np.stack([P.rhs(n, k) for k in range(4)], axis=1)); B.pr.apply_multi(X, False)
    print(case, "ok", float(np.abs(X).max()))
else:
    y = b.copy(); B.apply(y); z = b.copy(); B.apply_trans(z)
    if A is not None and case != "systolic" and case != "systolic_exact":
        L, U = oracle.port.ilu0(A); ref = oracle.port.apply_lu(L, U, b)
    else:
        ref = oracle.port.apply_llt(oracle.port.ichol0(A), b)
    print(case, "ok, relerr vs oracle %.2e" % (np.abs(y - ref).max() / np.abs(ref).max()), "part_mode", B.stats.get("part_mode"))
PY
for tool in ${TOOLS:-memcheck racecheck}; do
  for c in ${CASES:-systolic systolic_exact plane syncfree levels deep}; do
    echo "=== $tool $c"
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py $c 2>&1 | grep -v "^$" | tail -6
  done
done 2>&1 | tee gpurun_out/sanitizer_${TAG:-r2}.log
