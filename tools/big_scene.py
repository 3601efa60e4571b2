import importlib, sys
sys.path.insert(0,'.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
for n in (7000, 8192, 16384):
    sp,_ = host.generate_sweep(0xDEADBEEF, n); r.upload(sp, n)
    cam = host.main_camera(960, 16, seed=1)
    img, st = r.render(cam); img, st = r.render(cam)
    print(n, round(st.trace_ms,2), "ms", round(17*st.sphere_tests/st.trace_ms/1e9,2), "TFLOP/s")
