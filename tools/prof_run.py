"""Short single-GPU run for ncu: final scene, 1200x675 at a small spp (same kernel, fewer samples)."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200')
host = importlib.import_module('raytracing-with-zig_b200.host_api')
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
width = int(sys.argv[3]) if len(sys.argv) > 3 else 1200
sp, n = host.generate_world(0xDEADBEEF)
cam = host.main_camera(width, spp, seed=0xDEADBEEF)
r = pkg.Renderer(0)
r.upload(sp, n)
for _ in range(reps):
    img, st = r.render(cam)
print("trace_ms", st.trace_ms, "Msamples/s", st.samples / st.trace_ms / 1e3, "TFLOP/s", 17 * st.sphere_tests / st.trace_ms / 1e9)
