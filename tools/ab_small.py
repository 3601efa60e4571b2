"""A/B timing helper (development aid): C2 cameras and the C5 sweep up to 512 spheres, kernel ms only."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
def best(cam, reps=3):
    t = None
    for _ in range(reps):
        img, st = r.render(cam)
        t = st.trace_ms if t is None else min(t, st.trace_ms)
    return t, st
sp13, n13 = host.generate_chapter13(); r.upload(sp13, n13)
for name, kw in {"ch11": dict(look_from=(0, 0, 0), look_at=(0, 0, -1), vfov=90),
                 "ch13": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20, focus_dist=3.4, defocus_angle=10.0)}.items():
    t, st = best(host.camera_build(400, 16 / 9, spp=100, seed=0xDEADBEEF, **kw))
    print(f"C2 {name}: {t:.3f} ms  {st.samples / t / 1e3:.0f} Msamples/s")
for n in (16, 64, 256, 512):
    sp, _ = host.generate_sweep(0xDEADBEEF, n); r.upload(sp, n)
    t, st = best(host.main_camera(1920, 64, seed=0xDEADBEEF), 2)
    print(f"C5 N={n}: {t:.2f} ms  {17 * st.sphere_tests / t / 1e9:.2f} TFLOP/s")
