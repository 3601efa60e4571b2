"""Scratch GPU probe (development aid): times C3-like renders and the FP32 peak micro-benchmark."""
import ctypes as C, importlib, sys, time, json
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import rtzlib as R
pkg = importlib.import_module('raytracing-with-zig_b200')
l = pkg.lib()
v = C.c_double()
for variant in (0, 1, 2):
    rc = l.rtz_measure_fp32_peak(0, variant, C.byref(v)); print("fp32 peak variant", variant, rc, round(v.value, 2), "TFLOP/s", flush=True)
r = pkg.Renderer(0)
prng, sp, n = R.final_scene(0xDEADBEEF)
r.upload(sp, n)
for (w, spp) in [(400, 10), (1200, 50), (1200, 500), (1200, 500)]:
    cam = R.main_camera(w, spp, seed=0xDEADBEEF)
    img, st = r.render(cam)
    ms = st.trace_ms
    print(json.dumps(dict(w=w, spp=spp, trace_ms=round(ms, 3), resolve_ms=round(st.resolve_ms, 3), total_ms=round(st.total_ms, 3),
          msamples_s=round(st.samples / ms / 1e3, 1), gtests_s=round(st.sphere_tests / ms / 1e6, 1),
          tflops=round(17 * st.sphere_tests / ms / 1e9, 2), seg_per_sample=round(st.segments / st.samples, 4),
          capped=st.depth_capped, absorbed=st.absorbed)), flush=True)
sp13, n13 = R.chapter13_scene()
r.upload(sp13, n13)
cam = R.build_camera(400, 16/9, (-2, 2, 1), (0, 0, -1), 20, defocus=10.0, viewport_focus=3.4, focus=3.4, spp=100, seed=1)
img, st = r.render(cam); img, st = r.render(cam)
print("ch13", round(st.trace_ms, 3), "ms", round(st.samples / st.trace_ms / 1e3, 1), "Msamples/s seg/sample", st.segments / st.samples)
