for d in 0 1 2 3; do echo "== TPL_BLOCK_DBG=$d"; TPL_BLOCK_DBG=$d timeout 120 python scripts/blocked_probe.py --sizes 5000000,20000000 --k 40 --modes 5 2>&1 | grep "mode=5"; done
