cd /root/repo
timeout 200 python scripts/dev_tile.py parity 2>&1 | tail -1
timeout 300 python scripts/trace_tile.py p3d128 2>&1 | grep -E "apply|solve 0|tile duration|per tile|y lags|z lags" | head -8
timeout 300 python scripts/trace_tile.py p3d128 tile_warps=6 2>&1 | grep -E "apply|solve 0|tile duration|per tile|y lags|z lags" | head -8
