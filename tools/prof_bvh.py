"""Short single-GPU run of the BVH extension for ncu: final scene, <width> x ... at <spp>."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500; width = int(sys.argv[2]) if len(sys.argv) > 2 else 480
sp, n = host.generate_world(0xDEADBEEF)
cam = host.main_camera(width, spp, seed=0xDEADBEEF); cam.mode = 4
r = pkg.Renderer(0); r.upload(sp, n)
img, st = r.render(cam)
print("bvh trace_ms", st.trace_ms, "tests/segment", st.sphere_tests / st.segments)
