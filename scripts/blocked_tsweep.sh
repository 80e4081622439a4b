# tuning sweep of the blocked kernels: tile size x tile buffers (TPL_BLOCK_T / TPL_BLOCK_NTB, see tpl_blocks_host.h)
for NTB in 2 3 4; do for T in 1024 2048 3072; do echo "== NTB=$NTB T=$T"; TPL_BLOCK_NTB=$NTB TPL_BLOCK_T=$T timeout 300 python scripts/blocked_probe.py --sizes ${1:-5000000,20000000} --k 40 --modes 5 2>&1 | grep "^m=" | cut -c1-200; done; done
