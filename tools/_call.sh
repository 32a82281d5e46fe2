timeout 600 python tools/_dbg_bwd.py 2>&1 | tail -4
