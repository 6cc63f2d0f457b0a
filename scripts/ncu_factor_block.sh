cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cat > /tmp/fb_case.py <<'PY'
import sys
sys.path.insert(0, ".")
import torch, ilupp_b200 as ilupp
from ilupp_b200 import problems as P
ilupp.set_option("exact", 0)
A = P.irregular_powerlaw(n=150000)
B = ilupp.ILU0Preconditioner(A)
torch.cuda.synchronize()
print(B.timings_ms)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_factor_block -c 1 -f -o gpurun_out/r4_factor_block python /tmp/fb_case.py > gpurun_out/r4_factor_block.log 2>&1
tail -2 gpurun_out/r4_factor_block.log
