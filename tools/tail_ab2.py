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
K = ("RTZ_TAIL_WIDTH", "RTZ_TAIL_CHUNKS", "RTZ_COOP_MAX", "RTZ_PERM", "RTZ_CHUNK")
for name, env in (("image order, tail 4", {"RTZ_PERM": "0"}), ("scattered, tail 4", {}),
                  ("scattered, all across", {"RTZ_TAIL_CHUNKS": "100000"}), ("image order, all across", {"RTZ_PERM": "0", "RTZ_TAIL_CHUNKS": "100000"}),
                  ("scattered, all across w64", {"RTZ_TAIL_CHUNKS": "100000", "RTZ_TAIL_WIDTH": "64"}),
                  ("scattered, all across w16", {"RTZ_TAIL_CHUNKS": "100000", "RTZ_TAIL_WIDTH": "16"}),
                  ("scattered, chunk 64 tail 16", {"RTZ_CHUNK": "64", "RTZ_TAIL_CHUNKS": "16"}),
                  ("image order, chunk 64 tail 16", {"RTZ_PERM": "0", "RTZ_CHUNK": "64", "RTZ_TAIL_CHUNKS": "16"}),
                  ("scattered, tail 32", {"RTZ_TAIL_CHUNKS": "32"}), ("image order, tail 32", {"RTZ_PERM": "0", "RTZ_TAIL_CHUNKS": "32"})):
    for k in K:
        os.environ.pop(k, None)
    os.environ.update(env)
    print(f"{name:30s} shard 3/8: {best(sh):8.3f} ms   whole frame: {best(None, 2):8.3f} ms", flush=True)
