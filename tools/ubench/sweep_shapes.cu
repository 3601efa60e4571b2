// Micro-benchmark: the sphere-sweep loop in isolation, in several shapes (development aid).
//   packed (FFMA2, rays in pairs) vs scalar FFMA; 2 / 4 / 8 rays per thread; spheres from shared memory
//   (LDS.128 -> scalar-broadcast operands) vs constant bank (LDCU -> uniform-register operands).
// Reports algorithmic TFLOP/s at 17 FLOP per ray-sphere test and cycles per test per warp.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
constexpr int NS = 488;
struct ConstGeo { float4 geo[512]; };

template <int RAYS, bool PACKED, bool CONSTG>
__global__ void __launch_bounds__(128) k(const float4* __restrict__ geom, const __grid_constant__ ConstGeo C, int reps, unsigned* sink) {
    extern __shared__ float4 s_geo[];
    if (!CONSTG) {
        for (int i = threadIdx.x; i < NS; i += blockDim.x) s_geo[i] = geom[i];
        __syncthreads();
    }
    const float f = (float)(threadIdx.x + blockIdx.x * 7) * 1e-4f;
    float dx[RAYS], dy[RAYS], dz[RAYS], k1[RAYS], nk2[RAYS], tx[RAYS], ty[RAYS], tz[RAYS];
    unsigned m[RAYS];
#pragma unroll
    for (int r = 0; r < RAYS; ++r) {
        dx[r] = 0.6f + f * (r + 1), dy[r] = -0.3f + f * (r + 2), dz[r] = 0.2f - f * (r + 3);
        k1[r] = -1.0f + f * (r + 4), nk2[r] = -170.f - f * (r + 5);
        tx[r] = 26.f + f * (r + 6), ty[r] = 4.f - f * (r + 7), tz[r] = 6.f + f * (r + 8);
    }
    unsigned acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base < NS; base += 32) {
            const int cnt = min(32, NS - base);
#pragma unroll
            for (int r = 0; r < RAYS; ++r) m[r] = 0xFFFFFFFFu;
#pragma unroll 1
            for (int kk = 0; kk < cnt; kk += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 s = CONSTG ? C.geo[base + kk + u] : s_geo[base + kk + u];
                    if (PACKED) {
#pragma unroll
                        for (int r = 0; r < RAYS; r += 2) {
                            float2 h = __ffma2_rn(make_float2(dx[r], dx[r + 1]), make_float2(s.x, s.x), make_float2(k1[r], k1[r + 1]));
                            h = __ffma2_rn(make_float2(dy[r], dy[r + 1]), make_float2(s.y, s.y), h);
                            h = __ffma2_rn(make_float2(dz[r], dz[r + 1]), make_float2(s.z, s.z), h);
                            float2 w = __ffma2_rn(make_float2(tx[r], tx[r + 1]), make_float2(s.x, s.x), make_float2(nk2[r], nk2[r + 1]));
                            w = __ffma2_rn(make_float2(ty[r], ty[r + 1]), make_float2(s.y, s.y), w);
                            w = __ffma2_rn(make_float2(tz[r], tz[r + 1]), make_float2(s.z, s.z), w);
                            w = __fadd2_rn(w, make_float2(s.w, s.w));
                            const float2 d = __ffma2_rn(h, h, w);
                            m[r] = __funnelshift_l(__float_as_uint(d.x), m[r], 1);
                            m[r + 1] = __funnelshift_l(__float_as_uint(d.y), m[r + 1], 1);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < RAYS; ++r) {
                            float h = fmaf(dx[r], s.x, k1[r]);
                            h = fmaf(dy[r], s.y, h);
                            h = fmaf(dz[r], s.z, h);
                            float w = fmaf(tx[r], s.x, nk2[r]);
                            w = fmaf(ty[r], s.y, w);
                            w = fmaf(tz[r], s.z, w);
                            w = w + s.w;
                            const float d = fmaf(h, h, w);
                            m[r] = __funnelshift_l(__float_as_uint(d), m[r], 1);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RAYS; ++r) acc += __popc(~m[r]);
        }
        dx[0] += 1e-6f;
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}

template <int RAYS, bool PACKED, bool CONSTG>
void run(const char* name, const float4* d_geo, const ConstGeo& C, unsigned* sink, int sms, int threads_per_sm) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<RAYS, PACKED, CONSTG>, 128, CONSTG ? 0 : NS * 16);
    int want = threads_per_sm / 128;
    if (per_sm > want) per_sm = want;
    const int blocks = sms * per_sm, reps = 4000 / RAYS;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<RAYS, PACKED, CONSTG><<<blocks, 128, CONSTG ? 0 : NS * 16>>>(d_geo, C, reps, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<RAYS, PACKED, CONSTG>);
    const double tests = (double)NS * reps * RAYS * blocks * 128;
    const double tflops = tests * 17 / (best * 1e-3) / 1e12;
    const double cyc = best * 1e-3 * 1.965e9 * (sms * 4) / (tests / 32.0);
    printf("%-40s regs %3d  warps/SM %2d  %6.2f TFLOP/s (%4.1f%% of 74.45)  %5.2f cycles/test/warp\n", name, fa.numRegs,
           per_sm * 4, tflops, 100 * tflops / 74.45, cyc);
}

// Transposed packing: one FFMA2 = ONE ray against TWO spheres.  The ray constants are scalar-broadcast
// operands, the sphere pair is the packed operand (shared memory pair layout or constant bank).
struct ConstGeoT { float4 geo[512]; };   // [2*pair] = {cxA,cxB,cyA,cyB}, [2*pair+1] = {czA,czB,wA,wB}
template <int RAYS, bool CONSTG>
__global__ void __launch_bounds__(128) kt(const float4* __restrict__ geom, const __grid_constant__ ConstGeoT C, int reps, unsigned* sink) {
    extern __shared__ float4 s_geo[];
    if (!CONSTG) {
        for (int i = threadIdx.x; i < NS; i += blockDim.x) s_geo[i] = geom[i];
        __syncthreads();
    }
    const float f = (float)(threadIdx.x + blockIdx.x * 7) * 1e-4f;
    float dx[RAYS], dy[RAYS], dz[RAYS], k1[RAYS], nk2[RAYS], tx[RAYS], ty[RAYS], tz[RAYS];
    unsigned m[RAYS];
#pragma unroll
    for (int r = 0; r < RAYS; ++r) {
        dx[r] = 0.6f + f * (r + 1), dy[r] = -0.3f + f * (r + 2), dz[r] = 0.2f - f * (r + 3);
        k1[r] = -1.0f + f * (r + 4), nk2[r] = -170.f - f * (r + 5);
        tx[r] = 26.f + f * (r + 6), ty[r] = 4.f - f * (r + 7), tz[r] = 6.f + f * (r + 8);
    }
    unsigned acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base < NS; base += 32) {
            const int cnt = min(32, NS - base);
#pragma unroll
            for (int r = 0; r < RAYS; ++r) m[r] = 0xFFFFFFFFu;
#pragma unroll 1
            for (int kk = 0; kk < cnt; kk += 8) {
#pragma unroll
                for (int u = 0; u < 8; u += 2) {
                    const int pi = base + kk + u;   // pair index * 2
                    const float4 p0 = CONSTG ? C.geo[pi] : s_geo[pi];
                    const float4 p1 = CONSTG ? C.geo[pi + 1] : s_geo[pi + 1];
                    const float2 cx = make_float2(p0.x, p0.y), cy = make_float2(p0.z, p0.w), cz = make_float2(p1.x, p1.y), cw = make_float2(p1.z, p1.w);
#pragma unroll
                    for (int r = 0; r < RAYS; ++r) {
                        float2 h = __ffma2_rn(make_float2(dx[r], dx[r]), cx, make_float2(k1[r], k1[r]));
                        h = __ffma2_rn(make_float2(dy[r], dy[r]), cy, h);
                        h = __ffma2_rn(make_float2(dz[r], dz[r]), cz, h);
                        float2 w = __ffma2_rn(make_float2(tx[r], tx[r]), cx, make_float2(nk2[r], nk2[r]));
                        w = __ffma2_rn(make_float2(ty[r], ty[r]), cy, w);
                        w = __ffma2_rn(make_float2(tz[r], tz[r]), cz, w);
                        w = __fadd2_rn(w, cw);
                        const float2 d = __ffma2_rn(h, h, w);
                        m[r] = __funnelshift_l(__float_as_uint(d.x), m[r], 1);
                        m[r] = __funnelshift_l(__float_as_uint(d.y), m[r], 1);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RAYS; ++r) acc += __popc(~m[r]);
        }
        dx[0] += 1e-6f;
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}
template <int RAYS, bool CONSTG>
void runt(const char* name, const float4* d_geo, const ConstGeoT& C, unsigned* sink, int sms, int threads_per_sm) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kt<RAYS, CONSTG>, 128, CONSTG ? 0 : NS * 16);
    int want = threads_per_sm / 128;
    if (per_sm > want) per_sm = want;
    const int blocks = sms * per_sm, reps = 4000 / RAYS;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kt<RAYS, CONSTG><<<blocks, 128, CONSTG ? 0 : NS * 16>>>(d_geo, C, reps, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kt<RAYS, CONSTG>);
    const double tests = (double)NS * reps * RAYS * blocks * 128;
    const double tflops = tests * 17 / (best * 1e-3) / 1e12;
    const double cyc = best * 1e-3 * 1.965e9 * (sms * 4) / (tests / 32.0);
    printf("%-40s regs %3d  warps/SM %2d  %6.2f TFLOP/s (%4.1f%% of 74.45)  %5.2f cycles/test/warp\n", name, fa.numRegs,
           per_sm * 4, tflops, 100 * tflops / 74.45, cyc);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    std::vector<float4> g(512);
    for (int i = 0; i < 512; ++i) g[i] = make_float4((float)(i % 22) - 11.f, 0.2f, (float)(i / 22) - 11.f, i < NS ? -((i % 22 - 11.f) * (i % 22 - 11.f) + 0.f) : -INFINITY);
    float4* d_geo;
    unsigned* sink;
    cudaMalloc(&d_geo, 512 * 16), cudaMalloc(&sink, 64);
    cudaMemcpy(d_geo, g.data(), 512 * 16, cudaMemcpyHostToDevice);
    static ConstGeo C;
    memcpy(C.geo, g.data(), 512 * 16);
    static ConstGeoT CT;
    memcpy(CT.geo, g.data(), 512 * 16);
    const int sms = p.multiProcessorCount;
    for (int tps : {1024, 640}) {
        printf("-- up to %d threads per SM\n", tps);
        run<2, true, false>("packed  2 rays  smem", d_geo, C, sink, sms, tps);
        run<4, true, false>("packed  4 rays  smem", d_geo, C, sink, sms, tps);
        run<2, true, true>("packed  2 rays  const", d_geo, C, sink, sms, tps);
        run<4, true, true>("packed  4 rays  const", d_geo, C, sink, sms, tps);
        run<2, false, false>("scalar  2 rays  smem", d_geo, C, sink, sms, tps);
        run<4, false, false>("scalar  4 rays  smem", d_geo, C, sink, sms, tps);
        run<2, false, true>("scalar  2 rays  const", d_geo, C, sink, sms, tps);
        run<4, false, true>("scalar  4 rays  const", d_geo, C, sink, sms, tps);
        run<1, false, true>("scalar  1 ray   const", d_geo, C, sink, sms, tps);
        run<8, false, true>("scalar  8 rays  const", d_geo, C, sink, sms, tps);
        run<8, true, true>("packed  8 rays  const", d_geo, C, sink, sms, tps);
        runt<1, false>("transposed 1 ray  smem", d_geo, CT, sink, sms, tps);
        runt<2, false>("transposed 2 rays smem", d_geo, CT, sink, sms, tps);
        runt<4, false>("transposed 4 rays smem", d_geo, CT, sink, sms, tps);
        runt<1, true>("transposed 1 ray  const", d_geo, CT, sink, sms, tps);
        runt<2, true>("transposed 2 rays const", d_geo, CT, sink, sms, tps);
        runt<4, true>("transposed 4 rays const", d_geo, CT, sink, sms, tps);
    }
    return 0;
}
