"""A/B of the lockstep kernel against the warp-level wavefront kernel (development aid): same bytes, same work
counters, kernel ms.  usage: wave_ab.py [variants, default "0,8,9"] [sphere counts, default "16,64,128,256,512"]"""
import importlib, os, sys
sys.path.insert(0, '.')
pkg = importlib.import_module('raytracing-with-zig_b200'); host = importlib.import_module('raytracing-with-zig_b200.host_api')
variants = (sys.argv[1] if len(sys.argv) > 1 else "0,8,9").split(",")
counts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "16,64,128,256,512").split(",") if x]
c3 = len(sys.argv) > 3
def best(r, cam, reps=3):
    t = None
    for _ in range(reps):
        img, st = r.render(cam)
        t = st.trace_ms if t is None else min(t, st.trace_ms)
    return t, st, img.cpu()
ref = {}
for v in variants:
    os.environ["RTZ_VARIANT"] = v.split(":")[0]
    for kv in v.split(":")[1:]:
        k, val = kv.split("=")
        os.environ[k] = val
    r = pkg.Renderer(0)
    row = []
    sp13, n13 = host.generate_chapter13(); r.upload(sp13, n13)
    cases = [("ch11", n13, sp13, host.camera_build(400, 16 / 9, spp=100, seed=0xDEADBEEF, look_from=(0, 0, 0), look_at=(0, 0, -1), vfov=90)),
             ("ch13", n13, sp13, host.camera_build(400, 16 / 9, spp=100, seed=0xDEADBEEF, look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20, focus_dist=3.4, defocus_angle=10.0))]
    for n in counts:
        sp, _ = host.generate_sweep(0xDEADBEEF, n)
        cases.append((f"N={n}", n, sp, host.main_camera(1920, 64, seed=0xDEADBEEF)))
    if c3:
        sp, n = host.generate_world(0xDEADBEEF)
        cases.append(("C3/10", n, sp, host.main_camera(1200, 50, seed=0xDEADBEEF)))
    for name, n, sp, cam in cases:
        r.upload(sp, n)
        t, st, img = best(r, cam)
        key = (st.samples, st.segments, st.depth_capped, st.absorbed)
        if name not in ref:
            ref[name] = (img, key)
            same = "ref"
        else:
            same = "same" if (bool((img == ref[name][0]).all()) and key == ref[name][1]) else "DIFFERENT"
        row.append(f"{name} {t:.3f} ms {st.samples / t / 1e3:.0f} Ms/s {17 * st.sphere_tests / t / 1e9 / 74.45 * 100:.1f}% [{same}]")
    r.close()
    for kv in v.split(":")[1:]:
        os.environ.pop(kv.split("=")[0])
    print(f"variant {v}: " + " | ".join(row), flush=True)
