"""C3 at FULL size (1200x675, 500 spp): the GPU frame against the f64 CPU restatement of the reference at matched spp,
next to that restatement's own seed-to-seed noise floor (two independent reference renders).  Measurement tooling
(uses the oracle): ~2 x 65 s of CPU on 16 threads.  Writes gpurun_out/r2_c3_fullsize_parity.json."""
import ctypes as C, importlib, json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import rtzlib as R
pkg = importlib.import_module("raytracing-with-zig_b200")
host = importlib.import_module("raytracing-with-zig_b200.host_api")
orc = R.oracle()
T = orc.orc_hardware_threads()
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
rmse = lambda a, b: float(np.sqrt(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)))
sp, n = host.generate_world(0xDEADBEEF)
cam = host.main_camera(1200, spp, seed=0xDEADBEEF)
r = pkg.Renderer(0); r.upload(sp, n)
img, st = r.render(cam); img, st = r.render(cam)
g = img.cpu().numpy()
ocam = R.Camera.from_buffer_copy(bytes(cam)); osp = (R.Sphere * n).from_buffer_copy(bytes(sp)[: n * C.sizeof(R.Sphere)])
H, W = int(cam.height), int(cam.width)
a = np.zeros((H, W, 3), np.uint8); b = np.zeros_like(a); ast = R.Stats()
t0 = time.time(); orc.orc_render_philox64(C.byref(ocam), osp, n, 1, T, u8(a), None, C.byref(ast)); ta = time.time() - t0
t0 = time.time(); orc.orc_render_philox64(C.byref(ocam), osp, n, 2, T, u8(b), None, None); tb = time.time() - t0
d = g.astype(np.float64) - a.astype(np.float64)
out = dict(config=f"C3 final scene 1200x675, {spp} spp, depth 50 (full size)", gpu_ms=round(st.total_ms, 3),
           rmse_gpu_vs_f64_ref=round(rmse(g, a), 4), rmse_gpu_vs_second_ref=round(rmse(g, b), 4),
           noise_floor_ref_vs_ref=round(rmse(a, b), 4),
           rmse_per_channel=[round(float(np.sqrt(np.mean(d[..., c] ** 2))), 4) for c in range(3)],
           floor_per_channel=[round(rmse(a[..., c], b[..., c]), 4) for c in range(3)],
           bias_per_channel=[round(float(d[..., c].mean()), 4) for c in range(3)],
           psnr_gpu_vs_ref_db=round(20 * np.log10(255.0 / rmse(g, a)), 2), psnr_ref_vs_ref_db=round(20 * np.log10(255.0 / rmse(a, b)), 2),
           gpu_seg_per_sample=round(st.segments / st.samples, 4), ref_seg_per_sample=round(ast.segments / ast.samples, 4),
           gpu_depth_capped_pct=round(100 * st.depth_capped / st.samples, 4), ref_depth_capped_pct=round(100 * ast.depth_capped / ast.samples, 4),
           cpu_threads=T, cpu_seconds=[round(ta, 1), round(tb, 1)], cpu_msamples_s=round(ast.samples / ta / 1e6, 3),
           gpu_msamples_s=round(st.samples / st.total_ms / 1e3, 1))
print(json.dumps(out, indent=1))
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r2_c3_fullsize_parity.json").write_text(json.dumps(out, indent=1))
