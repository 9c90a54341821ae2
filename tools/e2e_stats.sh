FFMP_HOST_IO_STATS=1 python tools/e2e_ab.py child 16 3 2>&1 | tail -4
FFMP_HOST_IO_STATS=1 FFMP_HOST_IO=2 python tools/e2e_ab.py child 16 3 2>&1 | tail -4
