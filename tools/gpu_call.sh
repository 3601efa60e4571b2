python tools/c3_fullsize_parity.py 500 2>&1 | tail -32
