// Micro-benchmark: the sphere sweep with the FP64 pipe working beside the FP32 pipe (sm_100a, B200: the
// FP64 pipe issues one DFMA per 2 cycles per SM sub-partition and co-issues with packed FFMA2).
// Per unit of 6 spheres: 2 sphere PAIRS are tested with packed FP32 (8 FFMA2-class ops per pair and ray) and
// 2 spheres with FP64 (8 DFMA-class ops per sphere and ray).  Reports cycles per ray-sphere test.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o hybrid hybrid.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
constexpr int UNITS = 81;  // 81 * 6 = 486 spheres
struct Geo {
    float4 g32[UNITS * 6];   // per unit: 2 pairs x {cxA,cxB,cyA,cyB},{czA,czB,wA,wB} (3 pairs in FP32-only mode)
    double4 g64[UNITS * 2];  // per unit: 2 spheres {cx,cy,cz,w}
};
template <int RAYS, int MODE>  // MODE 0: FP32 only (3 pairs per unit), 1: hybrid
__global__ void __launch_bounds__(128) k(const __grid_constant__ Geo C, int reps, unsigned* sink) {
    const float f = (float)(threadIdx.x + blockIdx.x * 7) * 1e-4f;
    float dx[RAYS], dy[RAYS], dz[RAYS], k1[RAYS], nk2[RAYS], tx[RAYS], ty[RAYS], tz[RAYS];
    double Dx[RAYS], Dy[RAYS], Dz[RAYS], K1[RAYS], NK2[RAYS], Tx[RAYS], Ty[RAYS], Tz[RAYS];
    unsigned m[RAYS];
#pragma unroll
    for (int r = 0; r < RAYS; ++r) {
        dx[r] = 0.6f + f * (r + 1), dy[r] = -0.3f + f * (r + 2), dz[r] = 0.2f - f * (r + 3);
        k1[r] = -1.0f + f * (r + 4), nk2[r] = -170.f - f * (r + 5);
        tx[r] = 26.f + f * (r + 6), ty[r] = 4.f - f * (r + 7), tz[r] = 6.f + f * (r + 8);
        Dx[r] = dx[r], Dy[r] = dy[r], Dz[r] = dz[r], K1[r] = k1[r], NK2[r] = nk2[r], Tx[r] = tx[r], Ty[r] = ty[r], Tz[r] = tz[r];
    }
    unsigned acc = 0;
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll
        for (int r = 0; r < RAYS; ++r) m[r] = 0xFFFFFFFFu;
#pragma unroll 1
        for (int u = 0; u < UNITS; ++u) {
#pragma unroll
            for (int q = 0; q < (MODE == 0 ? 3 : 2); ++q) {
                const int pi = (MODE == 0 ? u * 6 : u * 4) + q * 2;
                const float4 p0 = C.g32[pi], p1 = C.g32[pi + 1];
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
            if (MODE == 1) {
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const double4 s = C.g64[u * 2 + q];
#pragma unroll
                    for (int r = 0; r < RAYS; ++r) {
                        double h = fma(Dx[r], s.x, K1[r]);
                        h = fma(Dy[r], s.y, h);
                        h = fma(Dz[r], s.z, h);
                        double w = fma(Tx[r], s.x, NK2[r]);
                        w = fma(Ty[r], s.y, w);
                        w = fma(Tz[r], s.z, w);
                        w = w + s.w;
                        const double d = fma(h, h, w);
                        m[r] = __funnelshift_l((unsigned)__double2hiint(d), m[r], 1);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RAYS; ++r) acc += __popc(~m[r]);
        dx[0] += 1e-6f, Dx[0] += 1e-6;
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}
template <int RAYS, int MODE>
void run(const char* name, const Geo& C, unsigned* sink, int sms, int threads_per_sm) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<RAYS, MODE>, 128, 0);
    int want = threads_per_sm / 128;
    if (per_sm > want) per_sm = want;
    const int blocks = sms * per_sm, reps = 3000 / RAYS;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<RAYS, MODE><<<blocks, 128>>>(C, reps, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<RAYS, MODE>);
    const double tests = (double)UNITS * 6 * reps * RAYS * blocks * 128;
    const double tflops = tests * 17 / (best * 1e-3) / 1e12;
    const double cyc = best * 1e-3 * 1.965e9 * (sms * 4) / (tests / 32.0);
    printf("%-34s regs %3d warps/SM %2d  %6.2f TF alg (%5.1f%% of 74.45)  %5.2f cycles/test/warp\n", name, fa.numRegs, per_sm * 4,
           tflops, 100 * tflops / 74.45, cyc);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    static Geo G;
    for (int i = 0; i < UNITS * 6; ++i) G.g32[i] = make_float4((float)(i % 22) - 11.f, 0.2f, (float)(i / 22) - 11.f, -((i % 22 - 11.f) * (i % 22 - 11.f)));
    for (int i = 0; i < UNITS * 2; ++i) G.g64[i] = make_double4((double)(i % 22) - 11., 0.2, (double)(i / 22) - 11., -((i % 22 - 11.) * (i % 22 - 11.)));
    unsigned* sink;
    cudaMalloc(&sink, 64);
    const int sms = p.multiProcessorCount;
    for (int tps : {768, 512}) {
        printf("-- up to %d threads per SM\n", tps);
        run<2, 0>("FP32 only, 2 rays", G, sink, sms, tps);
        run<2, 1>("FP32 + FP64 hybrid, 2 rays", G, sink, sms, tps);
        run<4, 0>("FP32 only, 4 rays", G, sink, sms, tps);
        run<4, 1>("FP32 + FP64 hybrid, 4 rays", G, sink, sms, tps);
        run<1, 1>("FP32 + FP64 hybrid, 1 ray", G, sink, sms, tps);
    }
    return 0;
}
