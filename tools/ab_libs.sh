#!/bin/bash
# tools/ab_libs.sh <libA> <libB>: same-box A/B of two builds on C2 / C5-small / C3 (kernel ms)
for i in 1 2; do for L in "$@"; do echo "== $L"; RTZ_LIB=$L python tools/ab_small.py | tr '\n' ';'; echo; RTZ_LIB=$L python tools/prof_run.py 500 3 1200; done; done
