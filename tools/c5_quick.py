import importlib, sys, json
sys.path.insert(0,'.')
pkg = importlib.import_module("raytracing-with-zig_b200"); host = importlib.import_module("raytracing-with-zig_b200.host_api")
r = pkg.Renderer(0)
for n in (512, 1024, 2048, 4096, 7000):
    sp,_ = host.generate_sweep(0xDEADBEEF, n); r.upload(sp, n)
    cam = host.main_camera(1920, 64 if n <= 4096 else 16, seed=0xDEADBEEF)
    best=None
    for _ in range(2):
        img, st = r.render(cam)
        best = st.total_ms if best is None else min(best, st.total_ms)
    print(n, round(best,2), "ms", round(17*st.sphere_tests/best/1e9,2), "TFLOP/s", round(17*st.sphere_tests/best/1e9/74.45,4))
