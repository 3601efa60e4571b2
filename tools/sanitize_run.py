"""Tiny renders for compute-sanitizer: both kernels (constant-bank and TMA/shared-memory), sharded and whole."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
sp, n = host.generate_world(0xDEADBEEF); r.upload(sp, n)
cam = host.main_camera(96, 3, seed=1)
img, st = r.render(cam); print("const kernel", st.samples, st.segments)
img, st = r.render(cam, pkg.rtz_shard(1, 3, 16, 16)); print("const kernel shard", st.samples)
sp, n = host.generate_sweep(0xDEADBEEF, 1024); r.upload(sp, n)
img, st = r.render(cam); print("smem kernel", st.samples, st.segments)
sp13, n13 = host.generate_chapter13()
rgb, st = pkg.render_host(host.camera_build(64, 16/9, (-2, 2, 1), (0, 0, -1), 20, focus_dist=3.4, defocus_angle=10.0, spp=4, seed=2), sp13, n13)
print("host path", st.samples)
