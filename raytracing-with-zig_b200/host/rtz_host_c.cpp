// rtz_host_c.cpp — a thin extern "C" facade over rtz_host.hpp so the Python tests / bench can drive
// the C++ host mirror (Scene / CameraBuilder / main) without a C++ test runner.  Product code:
// links librtz.so, never the oracle.
#include "rtz_host.hpp"

using namespace rtz;

extern "C" {

// Scene.init(seed) + generateWorld(): flattened spheres in list order; returns the count.
uint64_t rtzh_scene_generate_world(uint64_t seed, int32_t has_seed, rtz_sphere* out, uint64_t cap) {
    Scene s = Scene::init(has_seed ? std::optional<uint64_t>(seed) : std::nullopt);
    s.generateWorld();
    const auto f = s.world.flat();
    for (uint64_t i = 0; i < f.size() && i < cap; ++i) out[i] = f[i];
    return f.size();
}

// BASELINE config 5: exactly n spheres of the generalised final scene (0 on failure).
uint64_t rtzh_scene_generate_sweep(uint64_t seed, uint64_t n, rtz_sphere* out) {
    Scene s = Scene::init(seed);
    if (!s.generateSweep(n)) return 0;
    const auto f = s.world.flat();
    for (uint64_t i = 0; i < f.size(); ++i) out[i] = f[i];
    return f.size();
}

uint64_t rtzh_scene_generate_chapter13(rtz_sphere* out, uint64_t cap) {
    Scene s = Scene::init(0);
    s.generateChapter13();
    const auto f = s.world.flat();
    for (uint64_t i = 0; i < f.size() && i < cap; ++i) out[i] = f[i];
    return f.size();
}

// Camera.builder(width, aspect) with the setters applied in the order main.zig applies them:
// setDefocusAngle, setFocusDist (skipped when focus_dist < 0 -> builder default 10), setViewport,
// setSamplesPerPixel, setBounceMax, setVUp.  seed is carried through an (empty) Scene.
int32_t rtzh_camera_build(uint64_t width, double aspect, const double look_from[3], const double look_at[3],
                          const double vup[3], double vfov, double focus_dist, double defocus_angle, uint64_t spp,
                          uint64_t bounce_max, uint64_t seed, int32_t has_seed, rtz_camera* out) {
    try {
        Scene scene = Scene::init(has_seed ? std::optional<uint64_t>(seed) : std::optional<uint64_t>(0));
        if (!has_seed) scene.seed.reset();
        CameraBuilder b = Camera::builder(width, aspect);
        b.setScene(scene).setDefocusAngle(defocus_angle);
        if (focus_dist >= 0) b.setFocusDist(focus_dist);
        b.setViewport({look_from[0], look_from[1], look_from[2]}, {look_at[0], look_at[1], look_at[2]}, vfov)
            .setSamplesPerPixel(spp)
            .setBounceMax(bounce_max)
            .setVUp({vup[0], vup[1], vup[2]});
        *out = b.build().flat();
        return RTZ_OK;
    } catch (const std::exception&) {
        return RTZ_ERR_BAD_ARG;  // divScalar by zero etc. (the reference panics)
    }
}

// main(): config -> Scene -> Camera -> render -> images/<fileName>.  Returns the rtz status.
int32_t rtzh_main(uint64_t img_width, uint64_t samples_per_pixel, const char* file_name, uint64_t seed,
                  int32_t has_seed, int32_t num_gpus, rtz_stats* stats) {
    config().numGpus = num_gpus;
    config().imgWidth = img_width;
    config().samplesPerPixel = samples_per_pixel;
    config().fileName = file_name ? file_name : "chapter14.ppm";
    config().seed = has_seed ? std::optional<uint64_t>(seed) : std::nullopt;
    try {
        mainRender(stats);
        return RTZ_OK;
    } catch (const RenderFailed& e) {
        return e.status;
    } catch (const std::exception&) {
        return RTZ_ERR_BAD_ARG;
    }
}

// single-ray API, evaluated on the GPU through the probes
int32_t rtzh_list_hit(const rtz_sphere* sp, uint64_t n, const double o[3], const double d[3], double tmin,
                      double tmax, rtz_hit* out) {
    try {
        HittableList w;
        for (uint64_t i = 0; i < n; ++i) {
            MaterialArgs a;
            a.albedo = {sp[i].albedo[0], sp[i].albedo[1], sp[i].albedo[2]};
            a.fuzz = sp[i].fuzz, a.refractionIndex = sp[i].refraction_index;
            w.add(Hittable::init(HittableType::sphere, {{sp[i].center[0], sp[i].center[1], sp[i].center[2]}, sp[i].radius,
                                                        Material::init((MaterialType)sp[i].mat_type, a)}));
        }
        const auto r = w.hit(Ray::init({o[0], o[1], o[2]}, {d[0], d[1], d[2]}), Interval::init(tmin, tmax));
        std::memset(out, 0, sizeof *out);
        out->hit = r.has_value();
        if (r) {
            out->t = r->t, out->front = r->front;
            out->point[0] = r->point.x, out->point[1] = r->point.y, out->point[2] = r->point.z;
            out->normal[0] = r->normal.x, out->normal[1] = r->normal.y, out->normal[2] = r->normal.z;
        }
        return RTZ_OK;
    } catch (const RenderFailed& e) {
        return e.status;
    }
}

// PPM.init + PPM.save (ASCII P3, src/ppm.zig:25-39) / PPM.saveBinary (:42-60) for a caller-filled pixel array
// (3 f64 per pixel, or NULL for the all-zero image PPM.init leaves behind); Color.toRgb runs on the device.
int32_t rtzh_ppm_save(const char* path, uint64_t width, uint64_t height, const double* pixels, int32_t binary) {
    try {
        PPM ppm = PPM::init(width, height);
        if (pixels)
            for (size_t i = 0; i < ppm.pixels.size(); ++i) ppm.pixels[i] = Color::init(pixels[3 * i], pixels[3 * i + 1], pixels[3 * i + 2]);
        if (binary) ppm.saveBinary(path);
        else ppm.save(path);
        ppm.deinit();
        return RTZ_OK;
    } catch (const RenderFailed& e) {
        return e.status;
    }
}

// Color.fromValue / toValue / fromRgb / toRgb (src/color.zig:30-80)
void rtzh_color_from_value(uint32_t value, double out[3]) {
    const Color c = Color::fromValue(value);
    out[0] = c.pixel.x, out[1] = c.pixel.y, out[2] = c.pixel.z;
}
void rtzh_color_from_rgb(uint8_t r, uint8_t g, uint8_t b, double out[3]) {
    const Color c = Color::fromRgb(RGB{r, g, b});
    out[0] = c.pixel.x, out[1] = c.pixel.y, out[2] = c.pixel.z;
}
int32_t rtzh_color_to_value(const double in[3], uint32_t* out) {
    try {
        *out = Color::init(in[0], in[1], in[2]).toValue();
        return RTZ_OK;
    } catch (const RenderFailed& e) {
        return e.status;
    }
}
int32_t rtzh_color_to_rgb(const double in[3], uint8_t out[3]) {
    try {
        const RGB c = Color::init(in[0], in[1], in[2]).toRgb();
        out[0] = c.r, out[1] = c.g, out[2] = c.b;
        return RTZ_OK;
    } catch (const RenderFailed& e) {
        return e.status;
    }
}

}  // extern "C"
