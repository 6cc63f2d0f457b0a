cd /root/repo
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "tile_systolic" 2>&1 | grep -E "assert|Error|error|^E " | head -12
for z in 0 1 2 3; do timeout 600 python bench.py --steps 20 --warmup 5 --no-batched --quick --zero-copy $z 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('zero_copy $z: e2e ms', d['e2e']['ms_per_step'])"; done
