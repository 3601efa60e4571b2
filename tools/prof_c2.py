"""Short single-GPU run of BASELINE config 2 (chapter-13 scene, 400x225, 100 spp) for ncu: prof_c2.py [ch11|ch12|ch13] [spp] [width]."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
which = sys.argv[1] if len(sys.argv) > 1 else "ch13"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 100
width = int(sys.argv[3]) if len(sys.argv) > 3 else 400
cams = {"ch11": dict(look_from=(0, 0, 0), look_at=(0, 0, -1), vfov=90),
        "ch12": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20),
        "ch13": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20, focus_dist=3.4, defocus_angle=10.0)}
r = pkg.Renderer(0)
sp, n = host.generate_chapter13(); r.upload(sp, n)
cam = host.camera_build(width, 16 / 9, spp=spp, seed=0xDEADBEEF, **cams[which])
for _ in range(2):
    img, st = r.render(cam)
print(which, round(st.trace_ms, 3), "ms", round(st.samples / st.trace_ms / 1e3, 1), "Msamples/s", "seg/sample", round(st.segments / st.samples, 3))
