cd /root/repo
timeout 200 python scripts/dev_tile_time.py "2048 8 4" "2048 16 8" "128 128 128" sys_tile=1 2>&1 | tail -3
timeout 200 python scripts/dev_tile_time.py "128 128 128" sys_tile=1 tile_nap_fwd=0 tile_nap_fetch=0 tile_nap_tma=0 2>&1 | tail -1
timeout 200 python scripts/dev_tile_time.py "128 128 128" sys_tile=0 2>&1 | tail -1
