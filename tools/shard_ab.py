"""Per-shard trace time of an 8-way (or N-way) split of C3 on ONE GPU, for a few schedules (development aid)."""
import importlib, os, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
r = pkg.Renderer(0)
sp, n = host.generate_world(0xDEADBEEF); r.upload(sp, n)
cam = host.main_camera(1200, 500, seed=0xDEADBEEF)
K = ("RTZ_CHUNK", "RTZ_DRAIN")
def best(shard, reps=3):
    t = None
    for _ in range(reps):
        img, st = r.render(cam, shard)
        t = st.trace_ms if t is None else min(t, st.trace_ms)
    return t
whole = best(None, 2)
print(f"whole frame {whole:.3f} ms -> ideal shard {whole / world:.3f} ms")
for c in ("256", "96"):
    os.environ["RTZ_CHUNK"] = c
    print(f"whole frame, chunk {c}: {best(None, 2):.3f} ms")
os.environ.pop("RTZ_CHUNK")
for name, env in (("round-1 schedule (warps finish their own paths)", {"RTZ_DRAIN": "0"}), ("drain kernel (default)", {}),
                  ("drain kernel, chunk 256", {"RTZ_CHUNK": "256"}), ("drain kernel, chunk 96", {"RTZ_CHUNK": "96"})):
    for k in K:
        os.environ.pop(k, None)
    os.environ.update(env)
    ts = [best(pkg.rtz_shard(k, world, 4, 4)) for k in range(world)]
    print(f"{name:50s} max {max(ts):7.3f}  mean {sum(ts) / world:7.3f}  | " + " ".join(f"{t:6.2f}" for t in ts), flush=True)
