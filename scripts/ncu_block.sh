cd /root/repo
mkdir -p gpurun_out
cat > /tmp/ncu_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import torch, ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", 0)
A = P.irregular_powerlaw(n=200000); n = A.shape[0]
B = ilupp.ILU0Preconditioner(A)
x = torch.from_numpy(P.rhs(n, 0)).cuda()
for _ in range(3):
    B.apply(x)
torch.cuda.synchronize()
PY
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_trsv_block -s 2 -c 1 -f -o gpurun_out/r10_block python /tmp/ncu_case.py > gpurun_out/r10_block.log 2>&1
tail -2 gpurun_out/r10_block.log
timeout 200 ncu -i gpurun_out/r10_block.ncu-rep --page raw --csv > gpurun_out/r10_block_raw.csv 2>/dev/null; wc -c gpurun_out/r10_block_raw.csv
