"""Where does the end of a frame go?  Per-warp {start, queue ran dry, done} stamps (RTZ_TIMELINE=1, %globaltimer)
of one 1/8 shard of C3 and of the whole frame, for a few schedules.  Development aid / evidence for DESIGN.md."""
import importlib, os, sys, tempfile
import numpy as np
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
r = pkg.Renderer(0)
sp, n = host.generate_world(0xDEADBEEF); r.upload(sp, n)
cam = host.main_camera(1200, 500, seed=0xDEADBEEF)
out = os.path.join(tempfile.mkdtemp(), "tl.bin")
os.environ["RTZ_TIMELINE"] = "1"; os.environ["RTZ_TIMELINE_OUT"] = out
def run(shard):
    r.render(cam, shard)
    img, st = r.render(cam, shard)
    t = np.fromfile(out, dtype=np.uint64).reshape(-1, 4).astype(np.int64)
    t = t[t[:, 2] > 0]
    t0 = t[:, 0].min()
    start, dry, done = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3
    end = done.max()
    dry = np.where(t[:, 1] > 0, dry, done)
    q = lambda a, p: float(np.percentile(a, p))
    return (f"trace {st.trace_ms:7.3f} ms | warps {len(t)} | last start +{start.max():6.1f} us | queue dry first {end - dry.min():7.1f} / median {end - q(dry, 50):7.1f} us before the end | "
            f"warps done before the end: 10% {end - q(done, 10):7.1f}  50% {end - q(done, 50):7.1f}  90% {end - q(done, 90):6.1f}  99% {end - q(done, 99):6.1f} us | "
            f"idle warp-time in the tail {np.sum(end - done) / len(t):7.1f} us per warp")
for name, env in (("round-1 schedule (warps finish their own paths)", {"RTZ_DRAIN": "0"}), ("drain kernel (default)", {})):
    for k in ("RTZ_DRAIN",):
        os.environ.pop(k, None)
    os.environ.update(env)
    print(f"{name:48s} shard 3/8 : {run(pkg.rtz_shard(3, 8, 4, 4))}", flush=True)
    print(f"{name:48s} whole     : {run(None)}", flush=True)
