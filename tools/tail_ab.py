"""A/B of the frame tail (development aid): one shard of an 8-way split of C3 and the whole frame, with and without
the small chunks at the end of the queue and the sphere-parallel drain.  Same bytes in every configuration."""
import importlib, os, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
sp, n = host.generate_world(0xDEADBEEF); r.upload(sp, n)
cam = host.main_camera(1200, 500, seed=0xDEADBEEF)
sh = pkg.rtz_shard(3, 8, 4, 4)
def best(shard, reps=3):
    t = None
    for _ in range(reps):
        img, st = r.render(cam, shard)
        t = st.trace_ms if t is None else min(t, st.trace_ms)
    return t
for name, env in (("round-1 schedule", {"RTZ_TAIL_WIDTH": "0", "RTZ_COOP_MAX": "0", "RTZ_ORDER": "0"}), ("tail across pixels", {"RTZ_COOP_MAX": "0"}),
                  ("coop drain", {"RTZ_TAIL_WIDTH": "0"}), ("tail + coop, image order", {"RTZ_ORDER": "0"}), ("glass first + tail + coop", {}), ("both, coop 4", {"RTZ_COOP_MAX": "4"}), ("both, coop 8", {"RTZ_COOP_MAX": "8"}),
                  ("both, coop 24", {"RTZ_COOP_MAX": "24"}), ("both, coop 40", {"RTZ_COOP_MAX": "40"}), ("both, width 8", {"RTZ_TAIL_WIDTH": "8"}), ("both, width 64", {"RTZ_TAIL_WIDTH": "64"}), ("both, 2 chunks/warp", {"RTZ_TAIL_CHUNKS": "2"}), ("both, 8 chunks/warp", {"RTZ_TAIL_CHUNKS": "8"})):
    for k in ("RTZ_TAIL_WIDTH", "RTZ_TAIL_CHUNKS", "RTZ_COOP_MAX", "RTZ_ORDER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(f"{name:22s} shard 1/8: {best(sh):8.3f} ms   whole frame: {best(None, 2):8.3f} ms", flush=True)
