// rtz_scene.cuh — Scene.init + Scene.generateWorld / generateChapter13 ON THE DEVICE.
//
// The reference builds its world on the host from ONE sequential PRNG (src/Scene.zig:23-46, 48-134): Zig's
// std.Random.DefaultPrng = Xoshiro256++ seeded through SplitMix64, doubles from Random.float(f64).  The
// scene is a function of that exact stream, so the device generator restates both — one thread, f64, the
// reference's draw order (three draws per grid cell BEFORE the (4,0.2,0) exclusion test, then 6 | 4 | 0 per
// material) — and must give the oracle's spheres bit for bit (tests/test_gpu_parity.py).  It exists so that
// a benchmark or a caller without the reference's PRNG needs no other scene source; it is not a hot path.
//
// Everything here is IEEE f64 (+, *, sqrt) with no contraction (-fmad=false), like Zig's strict floats.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtz.h"

namespace rtz {

// std.Random.DefaultPrng (un-vendored Zig std; algorithm as published by Blackman & Vigna, seeding as
// Zig's Xoshiro256.init: four SplitMix64 outputs).  Call sites: src/Scene.zig:29-38, src/util.zig:15-17.
struct DevXoshiro {
    uint64_t s[4];
    __device__ static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    __device__ void seed(uint64_t x) {
        for (int i = 0; i < 4; ++i) {
            x += 0x9e3779b97f4a7c15ULL;
            uint64_t z = x;
            z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
            z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
            s[i] = z ^ (z >> 31);
        }
    }
    __device__ uint64_t next() {
        const uint64_t r = rotl(s[0] + s[3], 23) + s[0];
        const uint64_t t = s[1] << 17;
        s[2] ^= s[0], s[3] ^= s[1], s[1] ^= s[2], s[0] ^= s[3];
        s[2] ^= t;
        s[3] = rotl(s[3], 45);
        return r;
    }
    // Random.float(f64) = util.randomDouble (src/util.zig:15-17): 52 mantissa bits from the word, the
    // exponent from the leading zeros of its top 12 bits, extended by whole words while they are zero.
    __device__ double uniform() {
        const uint64_t r = next();
        uint64_t lz = (uint64_t)__clzll((long long)r);
        if (lz >= 12) {
            lz = 12;
            for (;;) {
                const uint64_t a = (uint64_t)__clzll((long long)next());
                lz += a;
                if (a != 64) break;
                if (lz >= 1022) {
                    lz = 1022;
                    break;
                }
            }
        }
        return __longlong_as_double((long long)(((1022 - lz) << 52) | (r & 0xFFFFFFFFFFFFFULL)));
    }
};

__device__ inline rtz_sphere make_dev_sphere(double cx, double cy, double cz, double r, int32_t mat, double ar, double ag,
                                             double ab, double fuzz, double ior) {
    rtz_sphere s;
    s.center[0] = cx, s.center[1] = cy, s.center[2] = cz;
    s.radius = r < 0.0 ? 0.0 : r;  // Sphere.init clamps (src/sphere.zig:18-24)
    s.mat_type = mat, s.reserved = 0;
    s.albedo[0] = ar, s.albedo[1] = ag, s.albedo[2] = ab;  // MaterialArgs defaults (src/material.zig:119-124)
    s.fuzz = fuzz, s.refraction_index = ior;
    return s;
}

// The body of generateWorld (src/Scene.zig:48-134) for grid cells a, b in [lo, hi).  Spheres are stored
// in FINAL order: ground at 0; with `big_first` the three big spheres at 1..3 and the small ones from 4
// (the C5 layout, so that truncation keeps the big ones), otherwise the small ones from 1 and the big
// three after them (the reference's order).  At most `keep` spheres are stored, all are counted, and the
// PRNG is always advanced over the whole grid.
__device__ inline uint64_t generate_grid(DevXoshiro& g, int lo, int hi, bool big_first, rtz_sphere* out, uint64_t keep) {
    uint64_t small = 0;
    const uint64_t first_small = big_first ? 4 : 1;
    if (out && keep > 0) out[0] = make_dev_sphere(0, -1000, 0, 1000, RTZ_MAT_LAMBERTIAN, 0.5, 0.5, 0.5, 0, 1.0);
    for (int a = lo; a < hi; ++a) {
        const double xOffset = (double)a;
        for (int b = lo; b < hi; ++b) {
            const double zOffset = (double)b;
            const double chooseMat = g.uniform();               // :66
            const double cx = xOffset + 0.9 * g.uniform();      // :68
            const double cz = zOffset + 0.9 * g.uniform();      // :70
            const double ex = cx - 4.0, ey = 0.2 - 0.2, ez = cz - 0.0;
            const double len = sqrt((ex * ex + ey * ey) + ez * ez);  // Vec.len: @reduce(.Add) left to right
            if (!(len > 0.9)) continue;                         // :73, tested AFTER the three draws
            rtz_sphere s;
            if (chooseMat < 0.8) {                              // :80-86: albedo = random * random
                const double lx = g.uniform(), ly = g.uniform(), lz = g.uniform();
                const double rx = g.uniform(), ry = g.uniform(), rz = g.uniform();
                s = make_dev_sphere(cx, 0.2, cz, 0.2, RTZ_MAT_LAMBERTIAN, lx * rx, ly * ry, lz * rz, 0, 1.0);
            } else if (chooseMat < 0.95) {                      // :87-95: randomRange(0.5,1) and fuzz in [0,0.5)
                const double ar = 0.5 + (1.0 - 0.5) * g.uniform(), ag = 0.5 + (1.0 - 0.5) * g.uniform(),
                             ab = 0.5 + (1.0 - 0.5) * g.uniform();
                const double fuzz = 0.0 + (0.5 - 0.0) * g.uniform();
                s = make_dev_sphere(cx, 0.2, cz, 0.2, RTZ_MAT_METAL, ar, ag, ab, fuzz, 1.0);
            } else {                                            // :75-78 glass
                s = make_dev_sphere(cx, 0.2, cz, 0.2, RTZ_MAT_DIELECTRIC, 1, 1, 1, 0, 1.5);
            }
            const uint64_t at = first_small + small;
            if (out && at < keep) out[at] = s;
            ++small;
        }
    }
    const uint64_t total = small + 4;
    const uint64_t big = big_first ? 1 : 1 + small;             // :106-133
    if (out && big + 0 < keep) out[big + 0] = make_dev_sphere(0, 1, 0, 1, RTZ_MAT_DIELECTRIC, 1, 1, 1, 0, 1.5);
    if (out && big + 1 < keep) out[big + 1] = make_dev_sphere(-4, 1, 0, 1, RTZ_MAT_LAMBERTIAN, 0.4, 0.2, 0.1, 0, 1.0);
    if (out && big + 2 < keep) out[big + 2] = make_dev_sphere(4, 1, 0, 1, RTZ_MAT_METAL, 0.7, 0.6, 0.5, 0, 1.0);
    return total;
}

struct SceneGenOut {
    uint64_t count;     // spheres of the scene (0: could not be generated)
    uint64_t state[4];  // Scene.prng after generation: what Camera.render would start from
};

// kind = RTZ_SCENE_*; one thread.
__global__ void scene_kernel(int kind, uint64_t seed, uint64_t n, rtz_sphere* out, uint64_t cap, SceneGenOut* res) {
    DevXoshiro g;
    g.seed(seed);
    uint64_t count = 0;
    if (kind == RTZ_SCENE_FINAL) {
        // a in 0..22 with xOffset = a - 11 (src/Scene.zig:62-65)
        count = generate_grid(g, -11, 11, false, out, cap);
    } else if (kind == RTZ_SCENE_CHAPTER13) {  // src/Scene.zig:136-182, no random draws
        const rtz_sphere w[5] = {
            make_dev_sphere(0, -100.5, -1, 100, RTZ_MAT_LAMBERTIAN, 0.8, 0.8, 0.0, 0, 1.0),
            make_dev_sphere(0, 0, -1.2, 0.5, RTZ_MAT_LAMBERTIAN, 0.1, 0.2, 0.5, 0, 1.0),
            make_dev_sphere(-1, 0, -1, 0.5, RTZ_MAT_DIELECTRIC, 1, 1, 1, 0, 1.5),
            make_dev_sphere(-1, 0, -1, 0.4, RTZ_MAT_DIELECTRIC, 1, 1, 1, 0, 1.0 / 1.5),
            make_dev_sphere(1, 0, -1, 0.5, RTZ_MAT_METAL, 0.8, 0.6, 0.2, 1, 1.0),
        };
        for (uint64_t i = 0; i < 5 && i < cap; ++i) out[i] = w[i];
        count = 5;
    } else if (kind == RTZ_SCENE_SWEEP && n >= 4) {
        // BASELINE config 5: generateWorld on a G x G grid centred on the origin, G = ceil(sqrt(n-4)), grown
        // until enough cells survive the exclusion, then cut to exactly n spheres (ground, the three big
        // ones, the small ones in generation order).  Every attempt replays the same stream.
        int G = (int)ceil(sqrt((double)(n - 4)));
        if (G < 1) G = 1;
        const int lo = -(G / 2), hi = lo + G;
        const DevXoshiro start = g;
        for (int grow = 0; grow < 64; ++grow) {
            g = start;
            if (generate_grid(g, lo - grow, hi + grow, true, nullptr, 0) >= n) {
                g = start;
                generate_grid(g, lo - grow, hi + grow, true, out, n < cap ? n : cap);
                count = n;
                break;
            }
        }
    }
    res->count = count;
    for (int i = 0; i < 4; ++i) res->state[i] = g.s[i];
}

}  // namespace rtz
