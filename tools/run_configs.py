"""Run the BASELINE.json configs on one GPU and print the results table of BASELINE.md §5.

  C1 deterministic goldens (bytes)         C2 chapter-13 scene, three cameras, 400x225, 100 spp
  C3 final scene 1200x675, 500 spp         C5 sphere-count sweep 16..4096, 1920x1080, 64 spp
Parity columns use the oracle (tests/rtzlib.py): this script is measurement tooling, not product code.
"""
import ctypes as C, importlib, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import rtzlib as R
pkg = importlib.import_module("raytracing-with-zig_b200")
host = importlib.import_module("raytracing-with-zig_b200.host_api")
orc = R.oracle()
PEAK = 148 * 128 * 2 * 1.965e9 / 1e12
THREADS = orc.orc_hardware_threads()
out = {"peak_tflops_nominal": PEAK, "cpu_threads": THREADS, "rows": []}

def u8(a): return a.ctypes.data_as(C.POINTER(C.c_uint8))
def rmse(a, b): return float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))

def timed(r, cam, reps=3):
    best = None
    for _ in range(reps):
        img, st = r.render(cam)
        if best is None or st.total_ms < best[1].total_ms:
            best = (img, st)
    return best

def row(name, st, n, extra=None):
    ms = st.total_ms
    d = dict(config=name, spheres=n, ms=round(ms, 3), msamples_s=round(st.samples / ms / 1e3, 1),
             mtests_s=round(st.sphere_tests / ms / 1e3, 1), tflops=round(17 * st.sphere_tests / ms / 1e9, 2),
             frac=round(17 * st.sphere_tests / ms / 1e9 / PEAK, 4), seg_per_sample=round(st.segments / max(1, st.samples), 4),
             capped_pct=round(100 * st.depth_capped / max(1, st.samples), 4))
    if extra: d.update(extra)
    out["rows"].append(d); print(json.dumps(d), flush=True)

r = pkg.Renderer(0)

# ---- C1 -------------------------------------------------------------------------------------
for mode, name, spheres in [(1, "chapter4", []), (2, "chapter5", [((0, 0, -1), .5)]), (3, "chapter6", [((0, 0, -1), .5), ((0, -100.5, -1), 100)])]:
    cam = R.Camera(); orc.orc_camera_legacy(400, 16 / 9, mode, C.byref(cam))
    sp = R.sphere_array([R.make_sphere(c, rr, 0) for c, rr in spheres]) if spheres else (R.Sphere * 1)()
    rgb, st = pkg.render_host(cam, sp, len(spheres)); rgb, st = pkg.render_host(cam, sp, len(spheres))
    ok = rgb.tobytes() == R.read_ppm(R.GOLDEN / f"{name}.ppm")[2]
    row(f"C1 {name} 400x225 1spp (f64 legacy kernel)", st, len(spheres), {"bytes_equal_golden": ok})

# ---- C2 -------------------------------------------------------------------------------------
sp13, n13 = host.generate_chapter13()
r.upload(sp13, n13)
cams = {"ch11": dict(look_from=(0, 0, 0), look_at=(0, 0, -1), vfov=90),
        "ch12": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20),
        "ch13": dict(look_from=(-2, 2, 1), look_at=(0, 0, -1), vfov=20, focus_dist=3.4, defocus_angle=10.0)}
for k, kw in cams.items():
    cam = host.camera_build(400, 16 / 9, spp=100, seed=0xDEADBEEF, **kw)
    img, st = timed(r, cam)
    ocam = R.Camera.from_buffer_copy(bytes(cam)); osp = (R.Sphere * n13).from_buffer_copy(bytes(sp13))
    a = np.zeros((225, 400, 3), np.uint8); b = np.zeros_like(a); ast = R.Stats()
    orc.orc_render_philox64(C.byref(ocam), osp, n13, 1, THREADS, u8(a), None, C.byref(ast))
    orc.orc_render_philox64(C.byref(ocam), osp, n13, 2, THREADS, u8(b), None, None)
    g = img.cpu().numpy()
    row(f"C2 {k} chapter-13 scene 400x225 100spp", st, n13,
        {"rmse_vs_f64_ref": round(rmse(g, a), 3), "noise_floor_ref_vs_ref": round(rmse(a, b), 3),
         "bias": [round(float(x), 3) for x in (g.astype(float) - a.astype(float)).mean(axis=(0, 1))],
         "ref_seg_per_sample": round(ast.segments / ast.samples, 4)})

# ---- C3 -------------------------------------------------------------------------------------
spw, nw = host.generate_world(0xDEADBEEF)
r.upload(spw, nw)
cam = host.main_camera(1200, 500, seed=0xDEADBEEF)
img, st = timed(r, cam)
cam_s = host.main_camera(400, 400, seed=0xDEADBEEF)
gs, _ = r.render(cam_s)
ocam = R.Camera.from_buffer_copy(bytes(cam_s)); osp = (R.Sphere * nw).from_buffer_copy(bytes(spw)[: nw * C.sizeof(R.Sphere)])
a = np.zeros((225, 400, 3), np.uint8); b = np.zeros_like(a); ast = R.Stats()
t0 = time.time(); orc.orc_render_philox64(C.byref(ocam), osp, nw, 1, THREADS, u8(a), None, C.byref(ast)); cpu_s = time.time() - t0
orc.orc_render_philox64(C.byref(ocam), osp, nw, 2, THREADS, u8(b), None, None)
g = gs.cpu().numpy()
row("C3 final scene 1200x675 500spp", st, nw,
    {"parity_at_400x225_400spp": {"rmse_vs_f64_ref": round(rmse(g, a), 3), "noise_floor_ref_vs_ref": round(rmse(a, b), 3),
                                  "bias": [round(float(x), 3) for x in (g.astype(float) - a.astype(float)).mean(axis=(0, 1))],
                                  "ref_seg_per_sample": round(ast.segments / ast.samples, 4)},
     "cpu_ref_msamples_s": round(ast.samples / cpu_s / 1e6, 3), "cpu_threads": THREADS})

# ---- C5 -------------------------------------------------------------------------------------
for n in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096):
    sp, _ = host.generate_sweep(0xDEADBEEF, n)
    r.upload(sp, n)
    cam = host.main_camera(1920, 64, seed=0xDEADBEEF)
    img, st = timed(r, cam, reps=2)
    row(f"C5 sweep N={n} 1920x1080 64spp", st, n)
Path(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "configs_r2.json").write_text(json.dumps(out, indent=1))
