"""Short single-GPU run of the C5 sweep scene for ncu: prof_sweep.py <n_spheres> [spp] [width]."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
n = int(sys.argv[1]); spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16; width = int(sys.argv[3]) if len(sys.argv) > 3 else 960
r = pkg.Renderer(0)
sp, _ = host.generate_sweep(0xDEADBEEF, n); r.upload(sp, n)
cam = host.main_camera(width, spp, seed=0xDEADBEEF)
img, st = r.render(cam)
print(n, "spheres", round(st.trace_ms, 3), "ms", round(17 * st.sphere_tests / st.trace_ms / 1e9, 2), "TFLOP/s")
