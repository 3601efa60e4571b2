// rtz_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A from-scratch f64 restatement of the reference renderer's `Camera.render` path
// (AndrewJarrett/raytracing-with-zig).  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load this library; the product
// (librtz.so) never links, imports or calls it.
//
// Parity status: PINNED.  The reference cannot be built here (no zig toolchain), so the
// oracle is certified by the reference's own golden vectors instead:
//   * test-files/chapter14.ppm  (reference src/main.zig:41-55, build.zig:62-66) — the whole
//     path incl. Zig std's Xoshiro256++/SplitMix64 and Random.float(f64), byte for byte;
//   * the 485-object count for seed 0xabadcafe (src/Scene.zig:189-205);
//   * the camera known answers (src/camera.zig:516-528) and every unit-test KAT listed in
//     SURVEY.md §4.2;
//   * test-files/chapter4/5/6.ppm for the legacy deterministic modes, test-binary.ppm for the
//     P6 writer.
// tests/test_oracle_golden.py checks all of them against fixtures copied under tests/golden/.
//
// Third-party arithmetic that is NOT under /root/reference: the Zig standard library
// (>= 0.14.0, build.zig.zon:12; un-vendored).  The pieces used by the path are restated
// here from their published algorithms: std.Random.DefaultPrng = Xoshiro256++ seeded by
// SplitMix64, Random.float(f64) (52 mantissa bits + geometric exponent), std.math.pow(x,5)
// (binary exponentiation), degreesToRadians (x * pi/180 constant).
//
// Build: g++ -O2 -ffp-contract=off  (Zig's strict IEEE semantics: no FMA contraction,
// @reduce(.Add) on a 3-vector is (e0+e1)+e2).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../include/rtz.h"
#include "philox_ref.h"

namespace {

// ------------------------------------------------------------------------------------------
// Zig std RNG (un-vendored dependency; see header comment)
// ------------------------------------------------------------------------------------------
struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        s += 0x9e3779b97f4a7c15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
};

inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }

// std.Random.DefaultPrng (reference src/Scene.zig:14,29-38)
struct Xoshiro256pp {
    uint64_t s[4];
    uint64_t draws = 0;  // u64 outputs consumed (diagnostics: SURVEY §3.4 "13.35 draws/sample")
    explicit Xoshiro256pp(uint64_t seed) {
        SplitMix64 g{seed};
        for (auto& x : s) x = g.next();
    }
    uint64_t next() {
        ++draws;
        const uint64_t r = rotl64(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl64(s[3], 45);
        return r;
    }
    // Random.float(f64): 52 random mantissa bits, exponent from the leading-zero count of the
    // remaining 12 bits (extended with more words when all 12 are zero).
    double nextDouble() {
        const uint64_t r = next();
        uint64_t lz = r ? (uint64_t)__builtin_clzll(r) : 64;
        if (lz >= 12) {
            lz = 12;
            for (;;) {
                const uint64_t w = next();
                const uint64_t a = w ? (uint64_t)__builtin_clzll(w) : 64;
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
        std::memcpy(&d, &bits, 8);
        return d;
    }
};

// Counter-based stream for the "all cores" CPU baseline: reference f64 arithmetic, but every
// (pixel, sample) owns an independent Philox4x32-10 stream (what north_star's RNG change makes
// possible).  53-bit doubles from two 32-bit words.
struct PhiloxStream {
    uint32_t key[2];
    uint32_t ctr[4];
    uint32_t buf[4];
    int have = 0;
    uint64_t draws = 0;
    PhiloxStream(uint64_t seed, uint32_t pixel, uint32_t sample) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = pixel;
        ctr[1] = sample;
        ctr[2] = 0xFFFFFFFFu;  // stream id distinct from the device's per-bounce streams
        ctr[3] = 0;
    }
    uint32_t next32() {
        if (have == 0) {
            philox4x32_10(ctr, key, buf);
            ++ctr[3];
            have = 4;
        }
        return buf[4 - have--];
    }
    double nextDouble() {
        ++draws;
        const uint64_t hi = next32(), lo = next32();
        return (double)(((hi << 32) | lo) >> 11) * 0x1.0p-53;
    }
};

// ------------------------------------------------------------------------------------------
// Vec (reference src/vec.zig:22-128) — strict f64, no contraction
// ------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 mulScalar(V3 v, double s) { return {v.x * s, v.y * s, v.z * s}; }  // vec.zig:35-37
// vec.zig:39-45: multiply by the reciprocal (Q2).  The reference panics on 0; the oracle
// reports that as NaNs through the same arithmetic (1/0 = inf) and flags it.
bool g_div_by_zero = false;
inline V3 divScalar(V3 v, double s) {
    if (s == 0) g_div_by_zero = true;
    const double r = 1.0 / s;
    return {v.x * r, v.y * r, v.z * r};
}
inline double dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }  // vec.zig:114-116 (Q4)
inline double lenSquared(V3 v) { return (v.x * v.x + v.y * v.y) + v.z * v.z; }  // vec.zig:51-53
inline double len(V3 v) { return std::sqrt(lenSquared(v)); }                   // vec.zig:47-49
inline V3 unit(V3 v) { return divScalar(v, len(v)); }                          // vec.zig:126-128
inline V3 cross(V3 a, V3 b) {                                                  // vec.zig:118-124
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// vec.zig:26-29: all components < 1e-8, WITHOUT abs (Q1)
inline bool nearZero(V3 v) { return v.x < 1e-8 && v.y < 1e-8 && v.z < 1e-8; }
// vec.zig:103-105: v - (n * dot(v,n)) * 2
inline V3 reflect(V3 v, V3 n) { return v - mulScalar(mulScalar(n, dot(v, n)), 2); }
// vec.zig:107-112
inline V3 refract(V3 v, V3 n, double etaiOverEtat) {
    const double cosTheta = std::fmin(dot(-v, n), 1.0);
    const V3 rPerp = mulScalar(v + mulScalar(n, cosTheta), etaiOverEtat);
    const V3 rParallel = mulScalar(n, -std::sqrt(std::fabs(1.0 - lenSquared(rPerp))));
    return rPerp + rParallel;
}

template <class Rng>
inline double randomDouble(Rng& g) { return g.nextDouble(); }  // util.zig:15-17
template <class Rng>
inline double randomDoubleRange(double mn, double mx, Rng& g) {  // util.zig:20-22
    return mn + (mx - mn) * randomDouble(g);
}
template <class Rng>
inline V3 vecRandom(Rng& g) {  // vec.zig:55-61 (x, then y, then z)
    const double a = randomDouble(g), b = randomDouble(g), c = randomDouble(g);
    return {a, b, c};
}
template <class Rng>
inline V3 vecRandomRange(double mn, double mx, Rng& g) {  // vec.zig:63-69
    const double a = randomDoubleRange(mn, mx, g), b = randomDoubleRange(mn, mx, g),
                 c = randomDoubleRange(mn, mx, g);
    return {a, b, c};
}
template <class Rng>
inline V3 randomUnitVec(Rng& g) {  // vec.zig:71-80: rejection, TRUE division (Q3)
    for (;;) {
        const V3 p = vecRandomRange(-1, 1, g);
        const double l2 = lenSquared(p);
        if (1e-160 < l2 && l2 <= 1) {
            const double s = std::sqrt(l2);
            return {p.x / s, p.y / s, p.z / s};
        }
    }
}
template <class Rng>
inline V3 randomInUnitDisk(Rng& g) {  // vec.zig:82-92
    for (;;) {
        const double a = randomDoubleRange(-1, 1, g), b = randomDoubleRange(-1, 1, g);
        const V3 p{a, b, 0};
        if (lenSquared(p) < 1) return p;
    }
}

// ------------------------------------------------------------------------------------------
// Ray, Interval (reference src/ray.zig:7-17, src/interval.zig:6-48)
// ------------------------------------------------------------------------------------------
struct Ray {
    V3 orig, dir;
    V3 at(double t) const { return orig + mulScalar(dir, t); }
};
struct Interval {
    double mn, mx;
    bool surrounds(double x) const { return mn < x && x < mx; }  // strict (Q8)
    bool contains(double x) const { return mn <= x && x <= mx; }
    double clamp(double x) const { return x < mn ? mn : (x > mx ? mx : x); }
};

// ------------------------------------------------------------------------------------------
// Sphere.hit (src/sphere.zig:26-54), HittableList.hit (src/hittable.zig:64-77)
// ------------------------------------------------------------------------------------------
struct HitRec {
    V3 point, normal;
    double t;
    bool front;
    int index;
};

inline bool sphereHit(const rtz_sphere& s, const Ray& ray, Interval t, HitRec& rec) {
    const V3 center{s.center[0], s.center[1], s.center[2]};
    const V3 oc = center - ray.orig;
    const double a = lenSquared(ray.dir);
    const double h = dot(ray.dir, oc);
    const double c = lenSquared(oc) - s.radius * s.radius;
    const double disc = h * h - a * c;
    if (disc < 0) return false;
    const double sqrtd = std::sqrt(disc);
    double root = (h - sqrtd) / a;
    if (!t.surrounds(root)) {
        root = (h + sqrtd) / a;
        if (!t.surrounds(root)) return false;
    }
    rec.t = root;
    rec.point = ray.at(root);
    const V3 outward = divScalar(rec.point - center, s.radius);
    rec.front = dot(ray.dir, outward) < 0;
    rec.normal = rec.front ? outward : -outward;
    return true;
}

// The sweep reads a compact copy {center, radius} of the spheres (32 B instead of the 80 B ABI
// struct) — identical arithmetic, friendlier to the L1 cache; it is what keeps the CPU baseline honest.
struct Geo {
    double cx, cy, cz, r;
};
struct World {
    const rtz_sphere* sp;
    uint64_t n;
    std::vector<Geo> geo;
    World(const rtz_sphere* s, uint64_t cnt) : sp(s), n(cnt), geo(cnt) {
        for (uint64_t i = 0; i < cnt; ++i) geo[i] = Geo{s[i].center[0], s[i].center[1], s[i].center[2], s[i].radius};
    }
};

inline bool listHit(const World& w, const Ray& ray, Interval t, HitRec& out) {
    int best = -1;
    double closest = t.mx;
    const Geo* g = w.geo.data();
    const V3 o = ray.orig, d = ray.dir;
    for (uint64_t i = 0; i < w.n; ++i) {
        // Sphere.hit (src/sphere.zig:27-42), root only; the record is built once for the winner,
        // which is what the reference's last assignment `hitRecord = tempRecord` leaves behind
        const V3 oc{g[i].cx - o.x, g[i].cy - o.y, g[i].cz - o.z};
        const double a = lenSquared(d);
        const double h = dot(d, oc);
        const double c = lenSquared(oc) - g[i].r * g[i].r;
        const double disc = h * h - a * c;
        if (disc < 0) continue;
        const double sqrtd = std::sqrt(disc);
        double root = (h - sqrtd) / a;
        if (!(t.mn < root && root < closest)) {
            root = (h + sqrtd) / a;
            if (!(t.mn < root && root < closest)) continue;
        }
        closest = root;
        best = (int)i;
    }
    if (best < 0) return false;
    const bool ok = sphereHit(w.sp[best], ray, Interval{t.mn, t.mx}, out);  // recomputes the same root
    (void)ok;
    out.index = best;
    return true;
}

inline bool listHit(const rtz_sphere* sp, uint64_t n, const Ray& ray, Interval t, HitRec& out) {
    return listHit(World(sp, n), ray, t, out);
}

// ------------------------------------------------------------------------------------------
// Material.scatter (src/material.zig:27-39, 55-68, 82-110)
// ------------------------------------------------------------------------------------------
inline double pow5(double x) { return x * ((x * x) * (x * x)); }  // std.math.pow(f64, x, 5)
inline double reflectance(double cosv, double ri) {               // material.zig:106-110
    double r0 = (1 - ri) / (1 + ri);
    r0 *= r0;
    return r0 + (1 - r0) * pow5(1 - cosv);
}

template <class Rng>
inline bool scatter(const rtz_sphere& s, const Ray& ray, const HitRec& rec, Rng& g, Ray& out,
                    V3& atten) {
    switch (s.mat_type) {
        case RTZ_MAT_LAMBERTIAN: {
            V3 dir = rec.normal + randomUnitVec(g);
            if (nearZero(dir)) dir = rec.normal;
            out = Ray{rec.point, dir};
            atten = V3{s.albedo[0], s.albedo[1], s.albedo[2]};
            return true;
        }
        case RTZ_MAT_METAL: {
            // the unit-vector draw happens even for fuzz == 0 (Q5); operands evaluate left to right
            const V3 refl = unit(reflect(ray.dir, rec.normal));
            const V3 reflected = refl + mulScalar(randomUnitVec(g), s.fuzz);
            if (dot(reflected, rec.normal) > 0) {
                out = Ray{rec.point, reflected};
                atten = V3{s.albedo[0], s.albedo[1], s.albedo[2]};
                return true;
            }
            return false;
        }
        default: {  // dielectric
            const double ri = rec.front ? 1.0 / s.refraction_index : s.refraction_index;
            const V3 unitDir = unit(ray.dir);
            const double cosT = std::fmin(dot(-unitDir, rec.normal), 1.0);
            const double sinT = std::sqrt(1.0 - cosT * cosT);
            const bool cannotRefract = ri * sinT > 1.0;
            const double approx = reflectance(cosT, ri);
            // `or` short-circuits: no draw on total internal reflection (Q6)
            const V3 d = (cannotRefract || approx > randomDouble(g)) ? reflect(unitDir, rec.normal)
                                                                     : refract(unitDir, rec.normal, ri);
            out = Ray{rec.point, d};
            atten = V3{1, 1, 1};
            return true;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Camera.getRay / rayColor / render (src/camera.zig:123-215)
// ------------------------------------------------------------------------------------------
struct Counters {
    uint64_t samples = 0, segments = 0, capped = 0, absorbed = 0, draws = 0;
};
inline V3 v3(const double* p) { return {p[0], p[1], p[2]}; }

template <class Rng>
inline Ray getRay(const rtz_camera& c, uint64_t i, uint64_t j, Rng& g) {
    const double ox = randomDouble(g) - 0.5;  // sampleSquare, :203-209
    const double oy = randomDouble(g) - 0.5;
    const V3 pixelSample =
        v3(c.pixel0) + mulScalar(v3(c.du), (double)i + ox) + mulScalar(v3(c.dv), (double)j + oy);
    V3 origin = v3(c.center);
    if (!(c.defocus_angle <= 0)) {  // :191-194, :212-215
        const V3 p = randomInUnitDisk(g);
        origin = v3(c.center) + mulScalar(v3(c.defocus_disk_u), p.x) + mulScalar(v3(c.defocus_disk_v), p.y);
    }
    return Ray{origin, pixelSample - origin};
}

template <class Rng>
inline V3 rayColor(const rtz_camera& c, const World& w, Ray ray, Rng& g, Counters& k) {
    const rtz_sphere* sp = w.sp;
    V3 ret{1, 1, 1};
    const Interval iv{c.t_min, c.t_max};
    for (uint64_t bounces = 0; bounces < c.bounce_max; ++bounces) {
        HitRec rec{};
        ++k.segments;
        if (listHit(w, ray, iv, rec)) {
            Ray sc;
            V3 att;
            if (scatter(sp[rec.index], ray, rec, g, sc, att)) {
                ray = sc;
                ret = ret * att;
                continue;
            }
            ++k.absorbed;
            return V3{0, 0, 0};
        }
        const double a = 0.5 * (unit(ray.dir).y + 1.0);
        const V3 sky = mulScalar(V3{1, 1, 1}, 1.0 - a) + mulScalar(V3{0.5, 0.7, 1}, a);
        return ret * sky;
    }
    ++k.capped;
    return V3{0, 0, 0};
}

// Color.toRgb (src/color.zig:63-80)
inline uint8_t toByte(double linear) {
    const double g = linear > 0 ? std::sqrt(linear) : 0;
    const Interval iv{0.000, 0.999};
    return (uint8_t)(256 * iv.clamp(g));
}
// legacy quantiser of the chapter4-6 goldens: trunc(255.999 * c), no gamma
inline uint8_t toByteLegacy(double c) { return (uint8_t)(255.999 * c); }

void fillStats(rtz_stats* st, const Counters& k, uint64_t n, uint64_t seed) {
    if (!st) return;
    std::memset(st, 0, sizeof(*st));
    st->samples = k.samples;
    st->segments = k.segments;
    st->sphere_tests = k.segments * n;
    st->depth_capped = k.capped;
    st->absorbed = k.absorbed;
    st->seed_used = seed;
}

}  // namespace

// ==========================================================================================
// extern "C" surface used by tests/ and bench.py through ctypes
// ==========================================================================================
extern "C" {

// ---- PRNG handles (Scene.init, src/Scene.zig:23-46) ---------------------------------------
void* orc_prng_new(uint64_t seed) { return new Xoshiro256pp(seed); }
void orc_prng_free(void* p) { delete (Xoshiro256pp*)p; }
uint64_t orc_prng_next(void* p) { return ((Xoshiro256pp*)p)->next(); }
double orc_prng_float(void* p) { return ((Xoshiro256pp*)p)->nextDouble(); }
uint64_t orc_prng_draws(void* p) { return ((Xoshiro256pp*)p)->draws; }
void orc_prng_state(void* p, uint64_t out[4]) { std::memcpy(out, ((Xoshiro256pp*)p)->s, 32); }

// ---- Scene builders -----------------------------------------------------------------------
static rtz_sphere mkSphere(V3 c, double r, int mat, V3 albedo, double fuzz, double ior) {
    rtz_sphere s;
    std::memset(&s, 0, sizeof(s));
    s.center[0] = c.x, s.center[1] = c.y, s.center[2] = c.z;
    s.radius = std::fmax(0.0, r);  // Sphere.init, src/sphere.zig:18-24 (Q16)
    s.mat_type = mat;
    s.albedo[0] = albedo.x, s.albedo[1] = albedo.y, s.albedo[2] = albedo.z;
    s.fuzz = fuzz;
    s.refraction_index = ior;
    return s;
}
// MaterialArgs defaults (src/material.zig:119-124): albedo (1,1,1), fuzz 0, ior 1.0
static const V3 kDefAlbedo{1, 1, 1};

// Scene.generateWorld (src/Scene.zig:48-134).  grid = 22 reproduces the reference; returns
// the number of spheres written (<= cap), or the number that WOULD be written if out == NULL.
static uint64_t generateGrid(Xoshiro256pp& g, int lo, int hi, rtz_sphere* out, uint64_t cap) {
    std::vector<rtz_sphere> w;
    w.push_back(mkSphere({0, -1000, 0}, 1000, RTZ_MAT_LAMBERTIAN, {0.5, 0.5, 0.5}, 0, 1.0));
    for (int a = lo; a < hi; ++a) {
        const double xOffset = (double)a;
        for (int b = lo; b < hi; ++b) {
            const double zOffset = (double)b;
            const double chooseMat = randomDouble(g);
            const double cx = xOffset + 0.9 * randomDouble(g);
            const double cz = zOffset + 0.9 * randomDouble(g);
            const V3 center{cx, 0.2, cz};
            if (len(center - V3{4, 0.2, 0}) > 0.9) {  // tested AFTER the three draws (Q15)
                if (chooseMat < 0.8) {
                    const V3 l = vecRandom(g);
                    const V3 r = vecRandom(g);
                    w.push_back(mkSphere(center, 0.2, RTZ_MAT_LAMBERTIAN, l * r, 0, 1.0));
                } else if (chooseMat < 0.95) {
                    const V3 albedo = vecRandomRange(0.5, 1, g);
                    const double fuzz = randomDoubleRange(0, 0.5, g);
                    w.push_back(mkSphere(center, 0.2, RTZ_MAT_METAL, albedo, fuzz, 1.0));
                } else {
                    w.push_back(mkSphere(center, 0.2, RTZ_MAT_DIELECTRIC, kDefAlbedo, 0, 1.5));
                }
            }
        }
    }
    w.push_back(mkSphere({0, 1, 0}, 1, RTZ_MAT_DIELECTRIC, kDefAlbedo, 0, 1.5));
    w.push_back(mkSphere({-4, 1, 0}, 1, RTZ_MAT_LAMBERTIAN, {0.4, 0.2, 0.1}, 0, 1.0));
    w.push_back(mkSphere({4, 1, 0}, 1, RTZ_MAT_METAL, {0.7, 0.6, 0.5}, 0, 1.0));
    if (out)
        for (uint64_t i = 0; i < w.size() && i < cap; ++i) out[i] = w[i];
    return w.size();
}

uint64_t orc_generate_world(void* prng, rtz_sphere* out, uint64_t cap) {
    // a in 0..22 with xOffset = a - 11  (src/Scene.zig:62-65)
    return generateGrid(*(Xoshiro256pp*)prng, -11, 11, out, cap);
}

// Scene.generateChapter13 (src/Scene.zig:136-182)
uint64_t orc_generate_chapter13(rtz_sphere* out, uint64_t cap) {
    const rtz_sphere w[5] = {
        mkSphere({0, -100.5, -1}, 100, RTZ_MAT_LAMBERTIAN, {0.8, 0.8, 0.0}, 0, 1.0),
        mkSphere({0, 0, -1.2}, 0.5, RTZ_MAT_LAMBERTIAN, {0.1, 0.2, 0.5}, 0, 1.0),
        mkSphere({-1, 0, -1}, 0.5, RTZ_MAT_DIELECTRIC, kDefAlbedo, 0, 1.5),
        mkSphere({-1, 0, -1}, 0.4, RTZ_MAT_DIELECTRIC, kDefAlbedo, 0, 1.0 / 1.5),
        mkSphere({1, 0, -1}, 0.5, RTZ_MAT_METAL, {0.8, 0.6, 0.2}, 1, 1.0),
    };
    for (uint64_t i = 0; i < 5 && i < cap; ++i) out[i] = w[i];
    return 5;
}

// BASELINE config 5: generateWorld generalised to a G x G grid centred on the
// origin, G = ceil(sqrt(n-4)), then truncated to exactly n spheres: ground first, the small
// spheres in generation order, and the three big spheres LAST-kept (they are moved to
// positions 1..3 so truncation never drops them).
uint64_t orc_generate_sweep(void* prng, uint64_t n, rtz_sphere* out) {
    if (n < 4) return 0;
    int G = (int)std::ceil(std::sqrt((double)(n - 4)));
    if (G < 1) G = 1;
    const int lo = -(G / 2), hi = lo + G;
    for (int grow = 0; grow < 64; ++grow) {
        Xoshiro256pp g = *(Xoshiro256pp*)prng;  // same stream for every attempt
        const uint64_t total = generateGrid(g, lo - grow, hi + grow, nullptr, 0);
        if (total >= n) {
            std::vector<rtz_sphere> w(total);
            generateGrid(g = *(Xoshiro256pp*)prng, lo - grow, hi + grow, w.data(), total);
            out[0] = w[0];
            out[1] = w[total - 3], out[2] = w[total - 2], out[3] = w[total - 1];
            for (uint64_t i = 4; i < n; ++i) out[i] = w[i - 3];
            *(Xoshiro256pp*)prng = g;
            return n;
        }
    }
    return 0;
}

// ---- Image / Viewport / CameraBuilder (src/camera.zig:26-80, 233-346) ----------------------
uint64_t orc_image_height(uint64_t width, double ratio) {  // Image.init :33-40 (Q12)
    const uint64_t h = (uint64_t)((double)width / ratio);
    return h < 1 ? 1 : h;
}
static const double kRadPerDeg = 0.017453292519943295769236907684886127134428718885417;
void orc_viewport(uint64_t w, uint64_t h, double vfov, double focusDist, double* vw, double* vh) {
    const double theta = vfov * kRadPerDeg;  // std.math.degreesToRadians (Q14)
    const double hh = std::tan(theta / 2.0);
    *vh = 2 * hh * focusDist;
    *vw = *vh * ((double)w / (double)h);
}
// CameraBuilder.build (:300-345).  `viewport_focus_dist` is the focusDist in force when
// setViewport ran (Q13: Viewport.init reads the value set SO FAR, default 10); `focus_dist`
// is the final one used by build().  Returns 1 if a divScalar/unit divided by zero (the
// reference would panic).
int orc_camera_build(uint64_t width, double aspect, const double lookFrom[3], const double lookAt[3],
                     const double vUp[3], double vfov, double viewport_focus_dist, double focus_dist,
                     double defocus_angle, uint64_t spp, uint64_t bounce_max, rtz_camera* out) {
    g_div_by_zero = false;
    std::memset(out, 0, sizeof(*out));
    out->width = width;
    out->height = orc_image_height(width, aspect);
    double vpw, vph;
    orc_viewport(out->width, out->height, vfov, viewport_focus_dist, &vpw, &vph);
    const V3 center = v3(lookFrom);
    const V3 w = unit(v3(lookFrom) - v3(lookAt));
    const V3 u = unit(cross(v3(vUp), w));
    const V3 v = cross(w, u);
    const V3 vu = mulScalar(u, vpw);
    const V3 vv = mulScalar(-v, vph);
    const V3 du = divScalar(vu, (double)out->width);
    const V3 dv = divScalar(vv, (double)out->height);
    const V3 upperLeft = center - mulScalar(w, focus_dist) - divScalar(vu, 2) - divScalar(vv, 2);
    const V3 pixel0 = upperLeft + mulScalar(du + dv, 0.5);
    const double defocusRadius = focus_dist * std::tan((defocus_angle / 2.0) * kRadPerDeg);
    const V3 ddu = mulScalar(u, defocusRadius), ddv = mulScalar(v, defocusRadius);
    auto put = [](double* d, V3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
    put(out->center, center), put(out->pixel0, pixel0), put(out->du, du), put(out->dv, dv);
    put(out->defocus_disk_u, ddu), put(out->defocus_disk_v, ddv);
    out->defocus_angle = defocus_angle;
    out->samples_per_pixel = spp;
    out->bounce_max = bounce_max;
    out->pixel_samples_scale = 1.0 / (double)spp;  // setSamplesPerPixel :282-286
    out->t_min = 1e-3;                              // Scene.interval, src/Scene.zig:21
    out->t_max = std::numeric_limits<double>::infinity();
    out->mode = RTZ_MODE_PATH;
    return g_div_by_zero ? 1 : 0;
}

// Camera of the book-chapter-4..6 pipeline behind test-files/chapter{4,5,6}.ppm:
// origin camera looking down -z, focal length 1, viewport height 2, width 2*(W/H).
void orc_camera_legacy(uint64_t width, double aspect, int32_t mode, rtz_camera* out) {
    std::memset(out, 0, sizeof(*out));
    out->width = width;
    out->height = orc_image_height(width, aspect);
    const double vph = 2.0, vpw = vph * ((double)out->width / (double)out->height);
    const V3 center{0, 0, 0};
    const V3 vu{vpw, 0, 0}, vv{0, -vph, 0};
    const V3 du = divScalar(vu, (double)out->width), dv = divScalar(vv, (double)out->height);
    const V3 upperLeft = center - V3{0, 0, 1.0} - divScalar(vu, 2) - divScalar(vv, 2);
    const V3 pixel0 = upperLeft + mulScalar(du + dv, 0.5);
    auto put = [](double* d, V3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; };
    put(out->center, center), put(out->pixel0, pixel0), put(out->du, du), put(out->dv, dv);
    out->samples_per_pixel = 1;
    out->bounce_max = 1;
    out->pixel_samples_scale = 1.0;
    out->t_min = 0.0;
    out->t_max = std::numeric_limits<double>::infinity();
    out->mode = mode;
}

// ---- Renders ------------------------------------------------------------------------------
// Camera.render with the reference's ONE shared sequential PRNG (single thread by nature).
// `prng` must be the handle that already generated the scene, so the stream continues where
// Scene.generateWorld left it (reference src/Scene.zig:20; src/main.zig:19-35).
int orc_render_reference(const rtz_camera* cam, const rtz_sphere* sp, uint64_t n, void* prng,
                         uint8_t* rgb, double* linear, rtz_stats* st) {
    if (!cam || !sp || !prng || cam->mode != RTZ_MODE_PATH) return RTZ_ERR_BAD_ARG;
    Xoshiro256pp& g = *(Xoshiro256pp*)prng;
    Counters k;
    const World world(sp, n);
    const uint64_t W = cam->width, H = cam->height;
    for (uint64_t j = 0; j < H; ++j)
        for (uint64_t i = 0; i < W; ++i) {
            V3 px{0, 0, 0};
            for (uint64_t s = 0; s < cam->samples_per_pixel; ++s) {
                const Ray r = getRay(*cam, i, j, g);
                px = px + rayColor(*cam, world, r, g, k);
                ++k.samples;
            }
            const V3 avg = mulScalar(px, cam->pixel_samples_scale);
            const uint64_t o = 3 * (i + j * W);
            if (linear) linear[o] = avg.x, linear[o + 1] = avg.y, linear[o + 2] = avg.z;
            if (rgb) rgb[o] = toByte(avg.x), rgb[o + 1] = toByte(avg.y), rgb[o + 2] = toByte(avg.z);
        }
    fillStats(st, k, n, 0);
    return RTZ_OK;
}

// Same f64 arithmetic, but every (pixel, sample) draws from its own Philox stream, rows
// striped over `threads` host threads: the "all cores" CPU baseline of BASELINE.md §4.
int orc_render_philox64(const rtz_camera* cam, const rtz_sphere* sp, uint64_t n, uint64_t seed,
                        int threads, uint8_t* rgb, double* linear, rtz_stats* st) {
    if (!cam || !sp || cam->mode != RTZ_MODE_PATH) return RTZ_ERR_BAD_ARG;
    if (threads < 1) threads = 1;
    const uint64_t W = cam->width, H = cam->height;
    const World world(sp, n);
    std::vector<Counters> ks(threads);
    auto work = [&](int tid) {
        Counters& k = ks[tid];
        for (uint64_t j = tid; j < H; j += threads)
            for (uint64_t i = 0; i < W; ++i) {
                V3 px{0, 0, 0};
                for (uint64_t s = 0; s < cam->samples_per_pixel; ++s) {
                    PhiloxStream g(seed, (uint32_t)(i + j * W), (uint32_t)s);
                    const Ray r = getRay(*cam, i, j, g);
                    px = px + rayColor(*cam, world, r, g, k);
                    ++k.samples;
                    k.draws += g.draws;
                }
                const V3 avg = mulScalar(px, cam->pixel_samples_scale);
                const uint64_t o = 3 * (i + j * W);
                if (linear) linear[o] = avg.x, linear[o + 1] = avg.y, linear[o + 2] = avg.z;
                if (rgb) rgb[o] = toByte(avg.x), rgb[o + 1] = toByte(avg.y), rgb[o + 2] = toByte(avg.z);
            }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < threads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    Counters k;
    for (auto& x : ks) k.samples += x.samples, k.segments += x.segments, k.capped += x.capped, k.absorbed += x.absorbed;
    fillStats(st, k, n, seed);
    return RTZ_OK;
}

// Legacy deterministic pipelines of test-files/chapter4/5/6.ppm: one ray through
// the pixel centre, hit interval (t_min, t_max) = (0, inf), no gamma, trunc(255.999 c).
int orc_render_legacy(const rtz_camera* cam, const rtz_sphere* sp, uint64_t n, uint8_t* rgb,
                      double* linear, rtz_stats* st) {
    if (!cam || cam->mode < RTZ_MODE_LEGACY_SKY || cam->mode > RTZ_MODE_LEGACY_NORMAL) return RTZ_ERR_BAD_ARG;
    Counters k;
    const uint64_t W = cam->width, H = cam->height;
    const Interval iv{cam->t_min, cam->t_max};
    for (uint64_t j = 0; j < H; ++j)
        for (uint64_t i = 0; i < W; ++i) {
            const V3 pc = v3(cam->pixel0) + mulScalar(v3(cam->du), (double)i) + mulScalar(v3(cam->dv), (double)j);
            const Ray ray{v3(cam->center), pc - v3(cam->center)};
            V3 col;
            HitRec rec{};
            bool hit = false;
            if (cam->mode != RTZ_MODE_LEGACY_SKY && n > 0) {
                ++k.segments;
                hit = listHit(sp, n, ray, iv, rec);
            }
            if (hit && cam->mode == RTZ_MODE_LEGACY_FLAT) {
                col = V3{1, 0, 0};
            } else if (hit) {
                col = mulScalar(rec.normal + V3{1, 1, 1}, 0.5);
            } else {
                const double a = 0.5 * (unit(ray.dir).y + 1.0);
                col = mulScalar(V3{1, 1, 1}, 1.0 - a) + mulScalar(V3{0.5, 0.7, 1.0}, a);
            }
            ++k.samples;
            const uint64_t o = 3 * (i + j * W);
            if (linear) linear[o] = col.x, linear[o + 1] = col.y, linear[o + 2] = col.z;
            if (rgb) rgb[o] = toByteLegacy(col.x), rgb[o + 1] = toByteLegacy(col.y), rgb[o + 2] = toByteLegacy(col.z);
        }
    fillStats(st, k, n, 0);
    return RTZ_OK;
}

// ---- PPM writers (src/ppm.zig:25-60) --------------------------------------------------------
int orc_write_ppm(const char* path, uint64_t w, uint64_t h, const uint8_t* rgb) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return RTZ_ERR_IO;
    std::fprintf(f, "P6\n%llu %llu\n255\n", (unsigned long long)w, (unsigned long long)h);
    std::fwrite(rgb, 1, (size_t)(3 * w * h), f);
    std::fputc('\n', f);  // trailing newline (Q17)
    return std::fclose(f) == 0 ? RTZ_OK : RTZ_ERR_IO;
}
int orc_write_ppm_ascii(const char* path, uint64_t w, uint64_t h, const uint8_t* rgb) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return RTZ_ERR_IO;
    std::fprintf(f, "P3\n%llu %llu\n255\n", (unsigned long long)w, (unsigned long long)h);
    for (uint64_t p = 0; p < w * h; ++p) std::fprintf(f, "%u %u %u\n", rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]);
    return std::fclose(f) == 0 ? RTZ_OK : RTZ_ERR_IO;
}

// ---- unit-level entry points for the KATs of SURVEY §4.2 ------------------------------------
void orc_to_rgb(const double* linear, uint64_t n, uint8_t* out) {
    for (uint64_t i = 0; i < 3 * n; ++i) out[i] = toByte(linear[i]);
}
double orc_linear_to_gamma(double x) { return x > 0 ? std::sqrt(x) : 0; }

static void fillHit(rtz_hit* out, bool hit, const HitRec& r) {
    std::memset(out, 0, sizeof(*out));
    out->hit = hit;
    if (!hit) return;
    out->index = r.index, out->front = r.front, out->t = r.t;
    out->point[0] = r.point.x, out->point[1] = r.point.y, out->point[2] = r.point.z;
    out->normal[0] = r.normal.x, out->normal[1] = r.normal.y, out->normal[2] = r.normal.z;
}
void orc_sphere_hit(const rtz_sphere* s, const double o[3], const double d[3], double tmin, double tmax,
                    rtz_hit* out) {
    HitRec r;
    r.index = 0;
    const bool hit = sphereHit(*s, Ray{v3(o), v3(d)}, Interval{tmin, tmax}, r);
    fillHit(out, hit, r);
}
void orc_list_hit(const rtz_sphere* sp, uint64_t n, const double o[3], const double d[3], double tmin,
                  double tmax, rtz_hit* out) {
    HitRec r;
    const bool hit = listHit(sp, n, Ray{v3(o), v3(d)}, Interval{tmin, tmax}, r);
    fillHit(out, hit, r);
}
// Material.scatter with the sequential PRNG handle.
void orc_scatter(const rtz_sphere* s, const double o[3], const double d[3], const rtz_hit* rec, void* prng,
                 rtz_scatter* out) {
    HitRec r;
    r.point = v3(rec->point), r.normal = v3(rec->normal), r.t = rec->t, r.front = rec->front != 0, r.index = rec->index;
    Ray sc{};
    V3 att{};
    std::memset(out, 0, sizeof(*out));
    out->scattered = scatter(*s, Ray{v3(o), v3(d)}, r, *(Xoshiro256pp*)prng, sc, att);
    if (out->scattered) {
        out->origin[0] = sc.orig.x, out->origin[1] = sc.orig.y, out->origin[2] = sc.orig.z;
        out->direction[0] = sc.dir.x, out->direction[1] = sc.dir.y, out->direction[2] = sc.dir.z;
        out->attenuation[0] = att.x, out->attenuation[1] = att.y, out->attenuation[2] = att.z;
    }
}
static void put3(double* d, V3 s) { d[0] = s.x, d[1] = s.y, d[2] = s.z; }
void orc_vec_unit(const double v[3], double out[3]) { put3(out, unit(v3(v))); }
void orc_vec_cross(const double a[3], const double b[3], double out[3]) { put3(out, cross(v3(a), v3(b))); }
double orc_vec_dot(const double a[3], const double b[3]) { return dot(v3(a), v3(b)); }
double orc_vec_len(const double a[3]) { return len(v3(a)); }
int orc_vec_near_zero(const double a[3]) { return nearZero(v3(a)); }
void orc_vec_reflect(const double v[3], const double n[3], double out[3]) { put3(out, reflect(v3(v), v3(n))); }
void orc_vec_refract(const double v[3], const double n[3], double e, double out[3]) { put3(out, refract(v3(v), v3(n), e)); }
void orc_vec_div_scalar(const double v[3], double s, double out[3]) { put3(out, divScalar(v3(v), s)); }
void orc_random_unit_vec(void* prng, double out[3]) { put3(out, randomUnitVec(*(Xoshiro256pp*)prng)); }
void orc_random_in_unit_disk(void* prng, double out[3]) { put3(out, randomInUnitDisk(*(Xoshiro256pp*)prng)); }
double orc_random_double_range(double mn, double mx, void* prng) { return randomDoubleRange(mn, mx, *(Xoshiro256pp*)prng); }
int orc_interval_surrounds(double mn, double mx, double x) { return Interval{mn, mx}.surrounds(x); }
int orc_interval_contains(double mn, double mx, double x) { return Interval{mn, mx}.contains(x); }
double orc_interval_clamp(double mn, double mx, double x) { return Interval{mn, mx}.clamp(x); }
double orc_reflectance(double c, double ri) { return reflectance(c, ri); }
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }
int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
