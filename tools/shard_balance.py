"""Load balance of the interleaved-tile sharding: render the 8 shards of C3 one after the other on one GPU."""
import importlib, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
sp, n = host.generate_world(0xDEADBEEF); r.upload(sp, n)
cam = host.main_camera(1200, 500, seed=0xDEADBEEF)
whole, st = r.render(cam); whole, st = r.render(cam)
print("whole frame", round(st.trace_ms, 2), "ms")
for world in (8, 4, 2):
    for tile in ((32, 8), (8, 8), (4, 4), (8, 2), (2, 2), (16, 1)):
        ms = []
        for rank in range(world):
            img, s2 = r.render(cam, pkg.rtz_shard(rank, world, tile[0], tile[1]))
            ms.append(s2.total_ms)
        print(world, tile, "max", round(max(ms), 2), "mean", round(sum(ms) / len(ms), 2), "max/mean", round(max(ms) / (sum(ms) / len(ms)), 4),
              "ideal", round(st.total_ms / world, 2))
