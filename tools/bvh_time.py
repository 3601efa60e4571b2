"""EXTENSION timing (not a bench line): C3 and the C5 sweep in RTZ_MODE_PATH_BVH beside the brute-force mode."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
def best(cam, reps=2):
    t = None
    for _ in range(reps):
        img, st = r.render(cam)
        t = st.trace_ms if t is None else min(t, st.trace_ms)
    return t, st
for name, n, width, spp in (("C3", 0, 1200, 500), ("C5 N=512", 512, 1920, 64), ("C5 N=4096", 4096, 1920, 64), ("N=16384", 16384, 960, 16)):
    sp, cnt = host.generate_world(0xDEADBEEF) if n == 0 else host.generate_sweep(0xDEADBEEF, n)
    r.upload(sp, cnt)
    cam = host.main_camera(width, spp, seed=0xDEADBEEF)
    cam.mode = 0; tb, sb = best(cam)
    cam.mode = 4; tv, sv = best(cam)
    print(f"{name}: brute {tb:.2f} ms ({sb.samples / tb / 1e3:.0f} Msamples/s, {sb.sphere_tests / sb.segments:.0f} tests/segment) | "
          f"BVH {tv:.2f} ms ({sv.samples / tv / 1e3:.0f} Msamples/s, {sv.sphere_tests / sv.segments:.1f} tests/segment) | x{tb / tv:.2f}")
