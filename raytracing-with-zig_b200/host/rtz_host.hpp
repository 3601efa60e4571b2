// rtz_host.hpp — C++ mirror of the reference's host-side API above the C ABI (include/rtz.h).
//
// The reference is Zig and there is no zig toolchain in this image, so the host side that a
// user of the reference sees — Scene, Camera/CameraBuilder, Hittable/HittableList, Sphere,
// Material, Color/RGB, PPM, Vec, Ray, Interval, util — is mirrored here in C++ with the same
// names, argument meaning and error behaviour, and `Camera::render()` is the drop-in: it
// flattens the scene and calls `rtz_render` (hand-written sm_100a CUDA) instead of running the
// triple loop of reference src/camera.zig:128-140 on the CPU.  The single-ray entry points
// (`Sphere::hit`, `HittableList::hit`, `Material::scatter`) also execute on the GPU through the
// probe calls of the ABI: nothing in this header intersects or shades on the CPU, and every
// compute call throws `RenderFailed` when there is no CUDA device.
//
// Host arithmetic (scene generation, camera derivation) is f64 in the reference's exact
// operation order (compile with -ffp-contract=off), because it is pinned by the reference's
// known answers: 485 spheres for seeds 0xdeadbeef / 0xabadcafe (src/Scene.zig:189-205), and
// du/dv/pixel0 of src/camera.zig:516-528.
//
// The Zig glue a reference maintainer would use instead is in ../zig/ (INTEGRATION.md).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "rtz.h"

namespace rtz {

struct RenderFailed : std::runtime_error {  // Zig glue maps non-zero status to error.RenderFailed
    int32_t status;
    RenderFailed(int32_t s, const std::string& what) : std::runtime_error(what), status(s) {}
};
inline void check(int32_t status) {
    if (status != RTZ_OK) throw RenderFailed(status, std::string(rtz_strerror(status)) + ": " + rtz_last_error());
}

// ---------------------------------------------------------------------------------------------
// std.Random.DefaultPrng of Zig >= 0.14 (un-vendored dependency of the reference): Xoshiro256++
// seeded through SplitMix64, and Random.float(f64).  reference src/Scene.zig:14,29-38.
// ---------------------------------------------------------------------------------------------
class DefaultPrng {
public:
    explicit DefaultPrng(uint64_t seed) {
        uint64_t x = seed;
        for (auto& w : s_) {
            x += 0x9e3779b97f4a7c15ULL;
            uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            w = z ^ (z >> 31);
        }
    }
    static DefaultPrng init(uint64_t seed) { return DefaultPrng(seed); }
    uint64_t next() {
        const uint64_t r = rotl(s_[0] + s_[3], 23) + s_[0];
        const uint64_t t = s_[1] << 17;
        s_[2] ^= s_[0], s_[3] ^= s_[1], s_[1] ^= s_[2], s_[0] ^= s_[3];
        s_[2] ^= t;
        s_[3] = rotl(s_[3], 45);
        return r;
    }
    // random().float(f64): 52 mantissa bits; exponent = 1022 - (leading zeros of the top 12 bits,
    // continued into further words when they are all zero)
    double floatF64() {
        const uint64_t r = next();
        uint64_t lz = clz(r);
        if (lz >= 12) {
            lz = 12;
            for (;;) {
                const uint64_t a = clz(next());
                lz += a;
                if (a != 64) break;
                if (lz >= 1022) {
                    lz = 1022;
                    break;
                }
            }
        }
        const uint64_t bits = ((1022 - lz) << 52) | (r & 0xFFFFFFFFFFFFFULL);
        double d;
        std::memcpy(&d, &bits, sizeof d);
        return d;
    }

private:
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    static uint64_t clz(uint64_t x) { return x ? (uint64_t)__builtin_clzll(x) : 64; }
    uint64_t s_[4];
};

// ---------------------------------------------------------------------------------------------
// util (reference src/util.zig)
// ---------------------------------------------------------------------------------------------
namespace util {
constexpr double pi = 3.14159265358979323846264338327950288419716939937510;
inline double degToRad(double degrees) { return degrees * pi / 180.0; }          // :8-10
inline double randomDouble(DefaultPrng* prng) { return prng->floatF64(); }       // :15-17
inline double randomDoubleRange(double mn, double mx, DefaultPrng* prng) {       // :20-22
    return mn + (mx - mn) * randomDouble(prng);
}
// std.math.degreesToRadians, what camera.zig actually uses (:19)
inline double degreesToRadians(double deg) { return deg * 0.017453292519943295769236907684886127134428718885417; }
}  // namespace util

// ---------------------------------------------------------------------------------------------
// Vec (reference src/vec.zig)
// ---------------------------------------------------------------------------------------------
struct Vec3 {
    double x = 0, y = 0, z = 0;
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    bool operator==(const Vec3& o) const { return x == o.x && y == o.y && z == o.z; }
};
using Point3 = Vec3;
using Color3 = Vec3;
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, Vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline Vec3 operator-(Vec3 a) { return {-a.x, -a.y, -a.z}; }

struct Vec {
    static Vec3 init(double x, double y, double z) { return {x, y, z}; }
    static Vec3 zero() { return {0, 0, 0}; }
    static Vec3 splat(double s) { return {s, s, s}; }
    static bool nearZero(Vec3 v) { return v.x < 1e-8 && v.y < 1e-8 && v.z < 1e-8; }  // :26-29, no abs
    static Vec3 addScalar(Vec3 v, double s) { return v + splat(s); }
    static Vec3 mulScalar(Vec3 v, double s) { return v * splat(s); }
    static Vec3 divScalar(Vec3 v, double s) {  // :39-45: panics on zero, multiplies by 1/s
        if (s == 0) throw std::domain_error("Trying to divide by zero!");
        return v * splat(1.0 / s);
    }
    static double lenSquared(Vec3 v) { return (v.x * v.x + v.y * v.y) + v.z * v.z; }
    static double len(Vec3 v) { return std::sqrt(lenSquared(v)); }
    static double dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
    static Vec3 cross(Vec3 a, Vec3 b) {
        return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    }
    static Vec3 unit(Vec3 v) { return divScalar(v, len(v)); }
    static Vec3 reflect(Vec3 v, Vec3 n) { return v - mulScalar(n * splat(dot(v, n)), 2); }
    static Vec3 refract(Vec3 v, Vec3 n, double etaiOverEtat) {
        const double cosTheta = std::fmin(dot(-v, n), 1.0);
        const Vec3 rPerp = mulScalar(v + mulScalar(n, cosTheta), etaiOverEtat);
        const Vec3 rParallel = mulScalar(n, -std::sqrt(std::fabs(1.0 - lenSquared(rPerp))));
        return rPerp + rParallel;
    }
    static Vec3 random(DefaultPrng* g) {
        const double a = util::randomDouble(g), b = util::randomDouble(g), c = util::randomDouble(g);
        return {a, b, c};
    }
    static Vec3 randomRange(double mn, double mx, DefaultPrng* g) {
        const double a = util::randomDoubleRange(mn, mx, g), b = util::randomDoubleRange(mn, mx, g),
                     c = util::randomDoubleRange(mn, mx, g);
        return {a, b, c};
    }
    static Vec3 randomUnitVec(DefaultPrng* g) {  // :71-80
        for (;;) {
            const Vec3 p = randomRange(-1, 1, g);
            const double l2 = lenSquared(p);
            if (1e-160 < l2 && l2 <= 1) {
                const double s = std::sqrt(l2);
                return {p.x / s, p.y / s, p.z / s};
            }
        }
    }
    static Vec3 randomInUnitDisk(DefaultPrng* g) {  // :82-92
        for (;;) {
            const double a = util::randomDoubleRange(-1, 1, g), b = util::randomDoubleRange(-1, 1, g);
            const Vec3 p{a, b, 0};
            if (lenSquared(p) < 1) return p;
        }
    }
    static Vec3 randomOnHemisphere(Vec3 normal, DefaultPrng* g) {  // :94-101
        const Vec3 u = randomUnitVec(g);
        return dot(u, normal) > 0.0 ? u : -u;
    }
};

// ---------------------------------------------------------------------------------------------
// Ray, Interval (reference src/ray.zig, src/interval.zig)
// ---------------------------------------------------------------------------------------------
struct Ray {
    Point3 orig;
    Vec3 dir;
    static Ray init(Point3 o, Vec3 d) { return {o, d}; }
    Vec3 at(double t) const { return orig + dir * Vec::splat(t); }
};

struct Interval {
    double min = std::numeric_limits<double>::infinity();
    double max = -std::numeric_limits<double>::infinity();
    static Interval empty() { return {}; }
    static Interval universe() {
        return {-std::numeric_limits<double>::infinity(), std::numeric_limits<double>::infinity()};
    }
    static Interval init(double mn, double mx) { return {mn, mx}; }
    double size() const { return max - min; }
    bool contains(double x) const { return min <= x && x <= max; }
    bool surrounds(double x) const { return min < x && x < max; }
    double clamp(double x) const { return x < min ? min : (x > max ? max : x); }
};

// ---------------------------------------------------------------------------------------------
// Color / RGB (reference src/color.zig).  toRgb quantises ON THE DEVICE (K3's to_byte) so the
// host API and the render output can never disagree.
// ---------------------------------------------------------------------------------------------
struct RGB {
    uint8_t r = 0, g = 0, b = 0;
    bool operator==(const RGB& o) const { return r == o.r && g == o.g && b == o.b; }
};
struct Color {
    Color3 pixel;
    static Color init(double r, double g, double b) { return {{r, g, b}}; }
    static Color fromVec(Vec3 v) { return {v}; }
    Vec3 toVec() const { return pixel; }
    static Color fromValue(uint32_t value) {  // u24, :30-38
        return {{(double)((value & 0xff0000) >> 16) / 255.999, (double)((value & 0x00ff00) >> 8) / 255.999,
                 (double)(value & 0x0000ff) / 255.999}};
    }
    static Color fromRgb(RGB c) { return {{(double)c.r / 255.999, (double)c.g / 255.999, (double)c.b / 255.999}}; }
    RGB toRgb() const {  // :63-80
        const double lin[3] = {pixel.x, pixel.y, pixel.z};
        uint8_t out[3];
        check(rtz_probe_to_rgb(lin, 1, out));
        return {out[0], out[1], out[2]};
    }
    uint32_t toValue() const {
        const RGB c = toRgb();
        return ((uint32_t)c.r << 16) | ((uint32_t)c.g << 8) | c.b;
    }
};

// ---------------------------------------------------------------------------------------------
// PPM (reference src/ppm.zig).  The render path hands over device-packed bytes; `pixels` keeps
// the reference's f64 Color view for API users that fill it themselves.
// ---------------------------------------------------------------------------------------------
struct PPM {
    size_t width = 0, height = 0;
    std::vector<Color> pixels;
    static PPM init(size_t w, size_t h) {
        PPM p;
        p.width = w, p.height = h, p.pixels.resize(w * h);
        return p;
    }
    void deinit() { pixels.clear(), pixels.shrink_to_fit(); }
    std::vector<uint8_t> quantise() const {
        std::vector<double> lin(3 * pixels.size());
        for (size_t i = 0; i < pixels.size(); ++i)
            lin[3 * i] = pixels[i].pixel.x, lin[3 * i + 1] = pixels[i].pixel.y, lin[3 * i + 2] = pixels[i].pixel.z;
        std::vector<uint8_t> rgb(3 * pixels.size());
        if (!pixels.empty()) check(rtz_probe_to_rgb(lin.data(), pixels.size(), rgb.data()));
        return rgb;
    }
    void saveBinary(const std::string& filename) const {  // :42-60
        const auto rgb = quantise();
        check(rtz_write_ppm(filename.c_str(), width, height, rgb.data()));
    }
    void save(const std::string& filename) const {  // ASCII P3, :25-39
        const auto rgb = quantise();
        FILE* f = std::fopen(filename.c_str(), "wb");
        if (!f) throw RenderFailed(RTZ_ERR_IO, filename);
        std::fprintf(f, "P3\n%zu %zu\n255\n", width, height);
        for (size_t i = 0; i < pixels.size(); ++i) std::fprintf(f, "%u %u %u\n", rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
        std::fclose(f);
    }
};

// ---------------------------------------------------------------------------------------------
// Material (reference src/material.zig:113-152)
// ---------------------------------------------------------------------------------------------
enum class MaterialType : int32_t { lambertian = RTZ_MAT_LAMBERTIAN, metal = RTZ_MAT_METAL, dielectric = RTZ_MAT_DIELECTRIC };

struct MaterialArgs {  // :119-124
    Color3 albedo{1, 1, 1};
    double fuzz = 0;
    DefaultPrng* prng = nullptr;
    double refractionIndex = 1.0;
};

struct HitRecord;
struct Scatter {  // :11-14
    Ray scattered;
    Color3 attenuation;
};

struct Material {
    MaterialType type = MaterialType::lambertian;
    Color3 albedo{1, 1, 1};
    double fuzz = 0;
    double refractionIndex = 1.0;
    DefaultPrng* prng = nullptr;
    static Material init(MaterialType t, const MaterialArgs& a) {
        Material m;
        m.type = t, m.prng = a.prng;
        if (t != MaterialType::dielectric) m.albedo = a.albedo;
        if (t == MaterialType::metal) m.fuzz = a.fuzz;
        if (t == MaterialType::dielectric) m.refractionIndex = a.refractionIndex;
        return m;
    }
    // Material.scatter(ray, rec): evaluated on the GPU (rtz_probe_scatter).  The device RNG is
    // counter based; its key for this call is one draw of the shared host PRNG, which keeps the
    // "same seed -> same scatter" property the reference's tests rely on (:168-194).
    std::optional<Scatter> scatter(const Ray& ray, const HitRecord& rec) const;
};

// ---------------------------------------------------------------------------------------------
// Sphere, HitRecord, Hittable, HittableList (reference src/sphere.zig, src/hittable.zig)
// ---------------------------------------------------------------------------------------------
struct HitRecord {  // hittable.zig:14-20
    Point3 point;
    Vec3 normal;
    Material mat;
    double t = 0;
    bool front = false;
};

inline rtz_sphere flatten(Point3 c, double r, const Material& m) {
    rtz_sphere s;
    std::memset(&s, 0, sizeof s);
    s.center[0] = c.x, s.center[1] = c.y, s.center[2] = c.z;
    s.radius = r;
    s.mat_type = (int32_t)m.type;
    s.albedo[0] = m.albedo.x, s.albedo[1] = m.albedo.y, s.albedo[2] = m.albedo.z;
    s.fuzz = m.fuzz;
    s.refraction_index = m.refractionIndex;
    return s;
}

struct Sphere {
    Point3 center;
    double radius = 0;
    Material mat;
    static Sphere init(Point3 c, double r, const Material& m) { return {c, std::fmax(0.0, r), m}; }  // :18-24
    rtz_sphere flat() const { return flatten(center, radius, mat); }
    std::optional<HitRecord> hit(const Ray& ray, Interval t) const;  // on the GPU
};

enum class HittableType { sphere };
struct SphereArgs {
    Point3 center;
    double radius;
    Material mat;
};
struct Hittable {  // tagged union with the single variant .sphere (hittable.zig:22-40)
    Sphere sphere;
    static Hittable init(HittableType, const SphereArgs& a) { return {Sphere::init(a.center, a.radius, a.mat)}; }
    std::optional<HitRecord> hit(const Ray& ray, Interval t) const { return sphere.hit(ray, t); }
};

inline std::optional<HitRecord> probeHit(const std::vector<rtz_sphere>& flat, const std::vector<const Material*>& mats,
                                         const Ray& ray, Interval t) {
    if (flat.empty()) return std::nullopt;
    const double o[3] = {ray.orig.x, ray.orig.y, ray.orig.z}, d[3] = {ray.dir.x, ray.dir.y, ray.dir.z};
    rtz_hit h;
    check(rtz_probe_hit(flat.data(), flat.size(), o, d, t.min, t.max, &h));
    if (!h.hit) return std::nullopt;
    HitRecord r;
    r.point = {h.point[0], h.point[1], h.point[2]};
    r.normal = {h.normal[0], h.normal[1], h.normal[2]};
    r.mat = *mats[h.index];
    r.t = h.t, r.front = h.front != 0;
    return r;
}

inline std::optional<HitRecord> Sphere::hit(const Ray& ray, Interval t) const {
    return probeHit({flat()}, {&mat}, ray, t);
}

struct HittableList {
    std::vector<Hittable> objects;
    static HittableList init() { return {}; }
    void deinit() { clear(); }
    void clear() { objects.clear(), objects.shrink_to_fit(); }
    void add(const Hittable& h) { objects.push_back(h); }
    std::vector<rtz_sphere> flat() const {
        std::vector<rtz_sphere> f;
        f.reserve(objects.size());
        for (const auto& o : objects) f.push_back(o.sphere.flat());
        return f;
    }
    std::optional<HitRecord> hit(const Ray& ray, Interval t) const {  // hittable.zig:64-77, on the GPU
        std::vector<const Material*> mats;
        for (const auto& o : objects) mats.push_back(&o.sphere.mat);
        return probeHit(flat(), mats, ray, t);
    }
};

inline std::optional<Scatter> Material::scatter(const Ray& ray, const HitRecord& rec) const {
    // a unit sphere tangent to the recorded hit reproduces (point, normal) for the probe
    const Point3 c = rec.front ? rec.point - rec.normal : rec.point + rec.normal;
    const rtz_sphere s = flatten(c, 1.0, *this);
    const double o[3] = {ray.orig.x, ray.orig.y, ray.orig.z}, d[3] = {ray.dir.x, ray.dir.y, ray.dir.z};
    const uint64_t key = prng ? prng->next() : 0;
    rtz_scatter out;
    check(rtz_probe_scatter(&s, 1, 0, o, d, key, 0, 0, 0, &out));
    if (!out.scattered) return std::nullopt;
    Scatter sc;
    sc.scattered = Ray::init({out.origin[0], out.origin[1], out.origin[2]}, {out.direction[0], out.direction[1], out.direction[2]});
    sc.attenuation = {out.attenuation[0], out.attenuation[1], out.attenuation[2]};
    return sc;
}

// ---------------------------------------------------------------------------------------------
// Scene (reference src/Scene.zig)
// ---------------------------------------------------------------------------------------------
struct Scene {
    HittableList world;
    std::optional<uint64_t> seed;
    std::shared_ptr<DefaultPrng> prng;
    Interval interval = Interval::init(1e-3, std::numeric_limits<double>::infinity());  // :21

    static Scene init(std::optional<uint64_t> seed) {  // :23-46
        Scene s;
        s.seed = seed;
        uint64_t v = 0;
        if (seed) {
            v = *seed;
        } else {  // std.posix.getrandom
            FILE* f = std::fopen("/dev/urandom", "rb");
            if (!f || std::fread(&v, 1, sizeof v, f) != sizeof v) throw std::runtime_error("getrandom failed");
            std::fclose(f);
        }
        s.prng = std::make_shared<DefaultPrng>(v);
        return s;
    }
    void deinit() { world.deinit(); }

    void generateWorld() { generateGrid(0, 22, 11); }  // :48-134: a, b in 0..22, offsets a-11, b-11

    // BASELINE config 5: generateWorld generalised to a G x G grid centred on the
    // origin, G = ceil(sqrt(n-4)) (grown until enough cells survive the (4,.2,0) exclusion), then cut
    // to exactly n spheres: ground, the three big spheres, then the small ones in generation order.
    // Not part of the reference; it exists so the sphere-count sweep needs no other scene source.
    bool generateSweep(size_t n) {
        if (n < 4) return false;
        int G = (int)std::ceil(std::sqrt((double)(n - 4)));
        if (G < 1) G = 1;
        const DefaultPrng start = *prng;
        for (int grow = 0; grow < 64; ++grow) {
            *prng = start;  // every attempt replays the same stream
            world.clear();
            const int lo = -(G / 2) - grow, hi = -(G / 2) + G + grow;
            generateGrid(lo, hi, 0);
            const size_t total = world.objects.size();
            if (total >= n) {
                std::vector<Hittable> cut;
                cut.push_back(world.objects[0]);
                for (size_t k = total - 3; k < total; ++k) cut.push_back(world.objects[k]);
                for (size_t k = 1; cut.size() < n; ++k) cut.push_back(world.objects[k]);
                world.objects = cut;
                return true;
            }
        }
        return false;
    }

    // the body of generateWorld for cells a, b in [lo, hi) with centre offsets (a - shift, b - shift)
    void generateGrid(int lo, int hi, int shift) {  // draw order per grid cell is part of the contract (Q15)
        DefaultPrng* g = prng.get();
        world.add(Hittable::init(HittableType::sphere,
                                 {{0, -1000, 0}, 1000, Material::init(MaterialType::lambertian, {{0.5, 0.5, 0.5}, 0, g, 1.0})}));
        for (int a = lo; a < hi; ++a) {
            const double xOffset = (double)a - shift;
            for (int b = lo; b < hi; ++b) {
                const double zOffset = (double)b - shift;
                const double chooseMat = util::randomDouble(g);
                const double cx = xOffset + 0.9 * util::randomDouble(g);
                const double cz = zOffset + 0.9 * util::randomDouble(g);
                const Point3 center{cx, 0.2, cz};
                if (Vec::len(center - Point3{4, 0.2, 0}) > 0.9) {
                    MaterialArgs glassArgs;
                    glassArgs.refractionIndex = 1.5, glassArgs.prng = g;
                    Material m = Material::init(MaterialType::dielectric, glassArgs);
                    if (chooseMat < 0.8) {
                        const Vec3 l = Vec::random(g);
                        const Vec3 r = Vec::random(g);
                        m = Material::init(MaterialType::lambertian, {l * r, 0, g, 1.0});
                    } else if (chooseMat < 0.95) {
                        const Vec3 albedo = Vec::randomRange(0.5, 1, g);
                        const double fuzz = util::randomDoubleRange(0, 0.5, g);
                        m = Material::init(MaterialType::metal, {albedo, fuzz, g, 1.0});
                    }
                    world.add(Hittable::init(HittableType::sphere, {center, 0.2, m}));
                }
            }
        }
        MaterialArgs a1;
        a1.refractionIndex = 1.5, a1.prng = g;
        world.add(Hittable::init(HittableType::sphere, {{0, 1, 0}, 1, Material::init(MaterialType::dielectric, a1)}));
        world.add(Hittable::init(HittableType::sphere,
                                 {{-4, 1, 0}, 1, Material::init(MaterialType::lambertian, {{0.4, 0.2, 0.1}, 0, g, 1.0})}));
        world.add(Hittable::init(HittableType::sphere,
                                 {{4, 1, 0}, 1, Material::init(MaterialType::metal, {{0.7, 0.6, 0.5}, 0, g, 1.0})}));
    }

    void generateChapter13() {  // :136-182
        DefaultPrng* g = prng.get();
        auto lam = [&](Color3 c) { return Material::init(MaterialType::lambertian, {c, 0, g, 1.0}); };
        auto glass = [&](double ior) {
            MaterialArgs a;
            a.refractionIndex = ior, a.prng = g;
            return Material::init(MaterialType::dielectric, a);
        };
        world.add(Hittable::init(HittableType::sphere, {{0, -100.5, -1}, 100, lam({0.8, 0.8, 0.0})}));
        world.add(Hittable::init(HittableType::sphere, {{0, 0, -1.2}, 0.5, lam({0.1, 0.2, 0.5})}));
        world.add(Hittable::init(HittableType::sphere, {{-1, 0, -1}, 0.5, glass(1.5)}));
        world.add(Hittable::init(HittableType::sphere, {{-1, 0, -1}, 0.4, glass(1.0 / 1.5)}));
        world.add(Hittable::init(HittableType::sphere,
                                 {{1, 0, -1}, 0.5, Material::init(MaterialType::metal, {{0.8, 0.6, 0.2}, 1, g, 1.0})}));
    }
};

// ---------------------------------------------------------------------------------------------
// config (reference build.zig:16-25 -> @import("config"))
// ---------------------------------------------------------------------------------------------
struct Config {
    size_t imgWidth = 3840;
    size_t samplesPerPixel = 500;
    std::string fileName = "chapter14.ppm";
    std::optional<uint64_t> seed;
    int32_t numGpus = 1;   // NEW build option -DnumGpus=N (0 = every GPU of the box): Camera.render then goes through
                           // rtz_render_multi — one process, no launcher (SURVEY 8b `num_gpus`, 8e)
};
inline Config& config() {
    static Config c;
    return c;
}

// ---------------------------------------------------------------------------------------------
// Image, Viewport, Camera, CameraBuilder (reference src/camera.zig)
// ---------------------------------------------------------------------------------------------
struct Image {
    size_t width = 100, height = 100;
    static Image init(size_t width, double ratio) {  // :33-40
        const size_t h = (size_t)((double)width / ratio);
        return {width, h < 1 ? 1 : h};
    }
    double aspectRatio() const { return (double)width / (double)height; }
};

struct Viewport {
    double width = 0, height = 0, vFov = 0;
    static Viewport init(Image img, double vFov, double focusDist) {  // :61-72
        const double theta = util::degreesToRadians(vFov);
        const double h = std::tan(theta / 2.0);
        const double height = 2 * h * focusDist;
        const double width = height * ((double)img.width / (double)img.height);
        return {width, height, vFov};
    }
};

namespace defaults {  // :218-232
constexpr size_t samplesPerPixel = 100;
constexpr size_t bounceMax = 50;
constexpr double focusDist = 10;
constexpr double defocusAngle = 0;
inline const Point3 cameraCenter{0, 0, 0};
inline const Point3 lookFrom{0, 0, 0};
inline const Point3 lookAt{0, 0, -1};
inline const Vec3 vUp{0, 1, 0};
}  // namespace defaults

struct CameraBuilder;

struct Camera {
    Image image;
    Viewport viewport;
    Scene scene;
    Point3 center = defaults::cameraCenter;
    size_t samplesPerPixel = defaults::samplesPerPixel;
    double pixelSamplesScale = 1.0 / (double)defaults::samplesPerPixel;
    size_t bounceMax = defaults::bounceMax;
    Point3 lookFrom = defaults::lookFrom, lookAt = defaults::lookAt;
    Vec3 vUp = defaults::vUp;
    Vec3 u, v, w;
    double focusDist = defaults::focusDist;
    Vec3 defocusDiskU, defocusDiskV;
    double defocusAngle = defaults::defocusAngle;
    Vec3 du, dv;
    Point3 pixel0;

    static CameraBuilder builder(size_t width, double aspectRatio);
    void deinit() { scene.deinit(); }

    rtz_camera flat() const {
        rtz_camera c;
        std::memset(&c, 0, sizeof c);
        auto put = [](double* d, Vec3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
        c.width = image.width, c.height = image.height;
        put(c.center, center), put(c.pixel0, pixel0), put(c.du, du), put(c.dv, dv);
        put(c.defocus_disk_u, defocusDiskU), put(c.defocus_disk_v, defocusDiskV);
        c.defocus_angle = defocusAngle;
        c.samples_per_pixel = samplesPerPixel, c.bounce_max = bounceMax;
        c.pixel_samples_scale = pixelSamplesScale;
        c.t_min = scene.interval.min, c.t_max = scene.interval.max;
        c.has_seed = scene.seed.has_value() ? 1 : 0;
        c.seed = scene.seed.value_or(0);
        c.mode = RTZ_MODE_PATH;
        return c;
    }

    // Camera.render (:123-145): ONE C-ABI call replaces the row/column/sample loop nest; the file
    // is written exactly as PPM.saveBinary does ("images/" ++ config.fileName, :144).
    void render(rtz_stats* stats = nullptr) const {
        const rtz_camera c = flat();
        const std::vector<rtz_sphere> spheres = scene.world.flat();
        std::vector<uint8_t> rgb(3 * image.width * image.height);
        if (config().numGpus == 1)
            check(rtz_render(&c, spheres.data(), spheres.size(), rgb.data(), stats));
        else
            check(rtz_render_multi(&c, spheres.data(), spheres.size(), config().numGpus, rgb.data(), stats));
        check(rtz_write_ppm(("images/" + config().fileName).c_str(), image.width, image.height, rgb.data()));
    }
};

struct CameraBuilder {
    Image image;
    std::optional<Scene> scene;
    size_t samplesPerPixel = defaults::samplesPerPixel;
    size_t bounceMax = defaults::bounceMax;
    Point3 center = defaults::cameraCenter, lookFrom = defaults::lookFrom, lookAt = defaults::lookAt;
    Vec3 vUp = defaults::vUp;
    double defocusAngle = defaults::defocusAngle;
    double focusDist = defaults::focusDist;
    std::optional<Viewport> viewport;
    double pixelSamplesScale = 1.0 / (double)defaults::samplesPerPixel;

    CameraBuilder& setScene(const Scene& s) { return scene = s, *this; }
    CameraBuilder& setFocusDist(double f) { return focusDist = f, *this; }       // must precede setViewport
    CameraBuilder& setDefocusAngle(double a) { return defocusAngle = a, *this; }
    CameraBuilder& setViewport(Point3 from, Point3 at, double vFov) {            // :273-279 (Q13)
        center = from, lookFrom = from, lookAt = at;
        viewport = Viewport::init(image, vFov, focusDist);
        return *this;
    }
    CameraBuilder& setSamplesPerPixel(size_t n) {
        samplesPerPixel = n, pixelSamplesScale = 1.0 / (double)n;
        return *this;
    }
    CameraBuilder& setBounceMax(size_t n) { return bounceMax = n, *this; }
    CameraBuilder& setVUp(Vec3 up) { return vUp = up, *this; }

    Camera build() const {  // :300-345
        if (!viewport) throw std::logic_error("setViewport was not called");  // `.?` on null panics in the reference
        Camera c;
        c.scene = scene ? *scene : Scene::init(std::nullopt);
        const Vec3 w = Vec::unit(lookFrom - lookAt);
        const Vec3 u = Vec::unit(Vec::cross(vUp, w));
        const Vec3 v = Vec::cross(w, u);
        const Vec3 vu = Vec::mulScalar(u, viewport->width);
        const Vec3 vv = Vec::mulScalar(-v, viewport->height);
        const Vec3 du = Vec::divScalar(vu, (double)image.width);
        const Vec3 dv = Vec::divScalar(vv, (double)image.height);
        const Vec3 upperLeft = center - Vec::mulScalar(w, focusDist) - Vec::divScalar(vu, 2) - Vec::divScalar(vv, 2);
        const Vec3 pixel0 = upperLeft + Vec::mulScalar(du + dv, 0.5);
        const double defocusRadius = focusDist * std::tan(util::degreesToRadians(defocusAngle / 2.0));
        c.image = image, c.viewport = *viewport;
        c.samplesPerPixel = samplesPerPixel, c.pixelSamplesScale = pixelSamplesScale, c.bounceMax = bounceMax;
        c.center = center, c.lookFrom = lookFrom, c.lookAt = lookAt, c.vUp = vUp;
        c.u = u, c.v = v, c.w = w;
        c.focusDist = focusDist;
        c.defocusDiskU = Vec::mulScalar(u, defocusRadius), c.defocusDiskV = Vec::mulScalar(v, defocusRadius);
        c.defocusAngle = defocusAngle;
        c.du = du, c.dv = dv, c.pixel0 = pixel0;
        return c;
    }
};

inline CameraBuilder Camera::builder(size_t width, double aspectRatio) {  // :109-117
    CameraBuilder b;
    b.image = Image::init(width, aspectRatio);
    return b;
}

// main (reference src/main.zig:14-36)
inline void mainRender(rtz_stats* stats = nullptr) {
    Scene scene = Scene::init(config().seed);
    scene.generateWorld();
    const double aspectRatio = 16.0 / 9.0;
    Camera camera = Camera::builder(config().imgWidth, aspectRatio)
                        .setScene(scene)
                        .setDefocusAngle(0.6)
                        .setFocusDist(10)
                        .setViewport(Point3{13, 2, 3}, Point3{0, 0, 0}, 20)
                        .setSamplesPerPixel(config().samplesPerPixel)
                        .build();
    camera.render(stats);
    camera.deinit();
}

}  // namespace rtz
