# half-pipeline timings of the blocked kernels (timing only, results are wrong): TPL_BLOCK_DBG 1 = fold off, 2 = compute off, 3 = both
for d in 0 1 2 3; do echo "== TPL_BLOCK_DBG=$d"; TPL_BLOCK_DBG=$d timeout 200 python scripts/blocked_probe.py --sizes ${1:-5000000,20000000} --k 40 --modes 5 2>&1 | grep "mode=5" | cut -c1-230; done
