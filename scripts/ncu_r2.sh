#!/bin/bash
# round-2 ncu captures (full set) of the kernels added this round: one launch each (run under gpurun, 1 GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/ncu_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import torch, ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", 0)
if sys.argv[1] == "packed":
    A = P.poisson3d(128); n = A.shape[0]
    perm = np.random.default_rng(1).permutation(n); A = A[perm][:, perm].tocsr(); A.sort_indices()
    B = ilupp.IChol0Preconditioner(A)
else:
    A = P.irregular_powerlaw(n=200000); n = A.shape[0]
    B = ilupp.ILU0Preconditioner(A)
x = torch.from_numpy(P.rhs(n, 0)).cuda()
for _ in range(3):
    B.apply(x)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trsv_level_packed -s 50 -c 3 -f -o gpurun_out/r2_level_packed python /tmp/ncu_case.py packed > gpurun_out/r2_level_packed.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trsv_block -s 2 -c 1 -f -o gpurun_out/r2_block python /tmp/ncu_case.py block > gpurun_out/r2_block.log 2>&1
tail -2 gpurun_out/r2_level_packed.log gpurun_out/r2_block.log
