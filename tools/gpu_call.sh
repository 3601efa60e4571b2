for i in 1 2; do
for L in tools/ab/librtz_prev.so tools/ab/librtz_flat.so; do echo "== $L: $(RTZ_LIB=$L python tools/prof_run.py 500 2 1200)"; done; done
for L in tools/ab/librtz_prev.so tools/ab/librtz_flat.so; do echo "== $L"; RTZ_LIB=$L python tools/wave_ab.py "0" "16,128,256,512" | cut -c1-600; RTZ_LIB=$L python tools/c5_quick.py | tr '\n' ';'; echo; done
