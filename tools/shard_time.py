"""Time one rank's share of C3 under interleaved-tile sharding on a single GPU (development aid):
   usage: shard_time.py [world] -> trace ms of rank 0 of `world`, best of 3."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sp, n = host.generate_world(0xDEADBEEF)
cam = host.main_camera(1200, 500, seed=0xDEADBEEF)
r = pkg.Renderer(0); r.upload(sp, n)
sh = pkg.rtz_shard(0, world, 32, 8)
best = None
for _ in range(4):
    img, st = r.render(cam, sh)
    best = st.trace_ms if best is None else min(best, st.trace_ms)
print(f"world {world}: rank 0 trace {best:.3f} ms, {st.samples / best / 1e3:.1f} Msamples/s, ideal share of 158.7 ms = {158.7 / world:.3f} ms")
