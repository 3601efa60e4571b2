// Micro-benchmark: which instruction of the sweep loop costs what (development aid).
// The loop is the render kernel's: one ray x a PAIR of spheres per packed instruction, sphere pairs in
// the constant bank (LDCU.64 -> UR operands).  Variants drop or replace one instruction class at a time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o sweep_ops sweep_ops.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>
constexpr int NS = 488;
struct ConstGeoT { float4 geo[512]; };  // [2*pair] = {cxA,cxB,cyA,cyB}, [2*pair+1] = {czA,czB,wA,wB}

// V: 0 baseline | 1 no disc op | 2 no FADD2 | 3 one LOP3 instead of two SHF | 5 disc as two scalar FFMA
//    6 sixteen spheres per inner iteration | 7 h-chain only (3 packed) | 8 no SHF, no disc (w-chain + fadd -> xor)
template <int RAYS, int V>
__global__ void __launch_bounds__(128) kt(const __grid_constant__ ConstGeoT C, int reps, unsigned* sink) {
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
    constexpr int STEP = (V == 6) ? 16 : 8;
    for (int rep = 0; rep < reps; ++rep) {
        for (int base = 0; base < NS; base += 32) {
            const int cnt = min(32, NS - base);
#pragma unroll
            for (int r = 0; r < RAYS; ++r) m[r] = 0xFFFFFFFFu;
#pragma unroll 1
            for (int kk = 0; kk < cnt; kk += STEP) {
                unsigned f8[RAYS];
#pragma unroll
                for (int r = 0; r < RAYS; ++r) f8[r] = 0;
#pragma unroll
                for (int u = 0; u < STEP; u += 2) {
                    const int pi = base + kk + u;
                    const float4 p0 = C.geo[pi];
                    const float4 p1 = C.geo[pi + 1];
                    const float2 cx = make_float2(p0.x, p0.y), cy = make_float2(p0.z, p0.w), cz = make_float2(p1.x, p1.y),
                                 cw = make_float2(p1.z, p1.w);
#pragma unroll
                    for (int r = 0; r < RAYS; ++r) {
                        float2 h = __ffma2_rn(make_float2(dx[r], dx[r]), cx, make_float2(k1[r], k1[r]));
                        h = __ffma2_rn(make_float2(dy[r], dy[r]), cy, h);
                        h = __ffma2_rn(make_float2(dz[r], dz[r]), cz, h);
                        float2 d;
                        if (V == 9) {
                            float2 w = __ffma2_rn(make_float2(tx[r], tx[r]), cx, cw);
                            w = __ffma2_rn(make_float2(ty[r], ty[r]), cy, w);
                            w = __ffma2_rn(make_float2(tz[r], tz[r]), cz, w);
                            d = __ffma2_rn(h, h, w);
                            if (d.x >= nk2[r]) f8[r] |= (0x80u >> (u & 7));
                            if (d.y >= nk2[r]) f8[r] |= (0x40u >> (u & 7));
                            continue;
                        }
                        if (V == 7) {
                            d = h;
                        } else {
                            float2 w = __ffma2_rn(make_float2(tx[r], tx[r]), cx, make_float2(nk2[r], nk2[r]));
                            w = __ffma2_rn(make_float2(ty[r], ty[r]), cy, w);
                            w = __ffma2_rn(make_float2(tz[r], tz[r]), cz, w);
                            if (V != 2) w = __fadd2_rn(w, cw);
                            if (V == 1 || V == 8) {
                                d = make_float2(__uint_as_float(__float_as_uint(w.x) ^ __float_as_uint(h.x)),
                                                __uint_as_float(__float_as_uint(w.y) ^ __float_as_uint(h.y)));
                            } else if (V == 5) {
                                d = make_float2(fmaf(h.x, h.x, w.x), fmaf(h.y, h.y, w.y));
                            } else {
                                d = __ffma2_rn(h, h, w);
                            }
                        }
                        if (V == 3 || V == 8) {
                            m[r] ^= __float_as_uint(d.x) ^ __float_as_uint(d.y);
                        } else {
                            m[r] = __funnelshift_l(__float_as_uint(d.x), m[r], 1);
                            m[r] = __funnelshift_l(__float_as_uint(d.y), m[r], 1);
                        }
                    }
                }
                if (V == 9) {
#pragma unroll
                    for (int r = 0; r < RAYS; ++r) m[r] = (m[r] << 8) | f8[r];
                }
            }
#pragma unroll
            for (int r = 0; r < RAYS; ++r) acc += __popc(~m[r]);
        }
        dx[0] += 1e-6f;
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}

template <int RAYS, int V>
void runt(const char* name, const ConstGeoT& C, unsigned* sink, int sms, int threads_per_sm, int packed_per_pair) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kt<RAYS, V>, 128, 0);
    int want = threads_per_sm / 128;
    if (per_sm > want) per_sm = want;
    const int blocks = sms * per_sm, reps = 4000 / RAYS;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kt<RAYS, V><<<blocks, 128>>>(C, reps, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kt<RAYS, V>);
    const double tests = (double)NS * reps * RAYS * blocks * 128;
    const double tflops = tests * 17 / (best * 1e-3) / 1e12;
    const double cyc = best * 1e-3 * 1.965e9 * (sms * 4) / (tests / 32.0);  // cycles per warp-test per SMSP
    printf("%-44s regs %3d warps/SM %2d  %6.2f TF alg (%5.1f%%)  %5.2f cyc/test  %5.2f cyc/packed-op (%d per pair)\n", name,
           fa.numRegs, per_sm * 4, tflops, 100 * tflops / 74.45, cyc, cyc * 2 / packed_per_pair, packed_per_pair);
}

// ---- single-instruction patterns with the sweep's operand kinds --------------------------------
#define N_ACC 8
template <int P>
__global__ void __launch_bounds__(128, 5) kp(const float* __restrict__ in, int iters, float* out,
                                             const __grid_constant__ ConstGeoT C) {
    float2 acc[N_ACC], A[N_ACC];
    for (int i = 0; i < N_ACC; ++i) {
        acc[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 1]);
        A[i] = make_float2(in[threadIdx.x + 40 + i], in[threadIdx.x + 41 + i]);
    }
    float s[4] = {in[threadIdx.x + 200], in[threadIdx.x + 201], in[threadIdx.x + 202], in[threadIdx.x + 203]};
    for (int it = 0; it < iters; ++it) {
        const float4 g0 = C.geo[(it & 63) * 2], g1 = C.geo[(it & 63) * 2 + 1];
        const float2 u0 = make_float2(g0.x, g0.y), u1 = make_float2(g0.z, g0.w), u2 = make_float2(g1.x, g1.y),
                     u3 = make_float2(g1.z, g1.w);
#pragma unroll
        for (int i = 0; i < N_ACC; ++i) {
            const float2 u = (i & 3) == 0 ? u0 : (i & 3) == 1 ? u1 : (i & 3) == 2 ? u2 : u3;
            if (P == 8) acc[i] = __ffma2_rn(make_float2(s[i & 3], s[i & 3]), u, acc[i]);  // scalar x UR pair + pair
            if (P == 9) acc[i] = __fadd2_rn(acc[i], u);                                    // pair + UR pair
            if (P == 10) acc[i] = __ffma2_rn(A[i], A[i], acc[i]);                          // h*h + w
            if (P == 11) acc[i] = __ffma2_rn(make_float2(s[0], s[0]), u, acc[i]);          // one scalar reused
            if (P == 12) { acc[i].x = fmaf(s[i & 3], u.x, acc[i].x); acc[i].y = fmaf(s[i & 3], u.y, acc[i].y); }  // scalar FFMA R,UR,R
            if (P == 13) { acc[i].x = fmaf(A[i].x, A[i].x, acc[i].x); acc[i].y = fmaf(A[i].y, A[i].y, acc[i].y); }
        }
    }
    float r = 0;
    for (int i = 0; i < N_ACC; ++i) r += acc[i].x + acc[i].y;
    if (r == 1.2345f) out[0] = r;
}
template <int P>
void runp(const char* name, const float* in, float* out, const ConstGeoT& C, int sms) {
    const int iters = 1 << 15, blocks = sms * 5, threads = 128;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kp<P><<<blocks, threads>>>(in, iters, out, C);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    const bool scalar = (P == 12 || P == 13);
    const double slots = (double)iters * N_ACC * blocks * threads;
    const double warp_instr = slots / 32.0 * (scalar ? 2 : 1);
    const double cyc = best * 1e-3 * 1.965e9 * (sms * 4) / warp_instr;
    printf("%-52s %5.2f cycles per warp-instruction per SMSP\n", name, cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    std::vector<float4> g(512);
    for (int i = 0; i < 512; ++i)
        g[i] = make_float4((float)(i % 22) - 11.f, 0.2f, (float)(i / 22) - 11.f,
                           i < NS ? -((i % 22 - 11.f) * (i % 22 - 11.f) + 0.f) : -INFINITY);
    unsigned* sink;
    float *in, *out;
    cudaMalloc(&sink, 64), cudaMalloc(&in, 4096), cudaMalloc(&out, 64);
    cudaMemset(in, 0, 4096);
    static ConstGeoT CT;
    memcpy(CT.geo, g.data(), 512 * 16);
    const int sms = p.multiProcessorCount;
    for (int tps : {2048, 1536, 1024, 768}) {
        printf("-- up to %d threads per SM\n", tps);
        runt<2, 0>("V0 baseline (8 packed + 2 SHF per pair)", CT, sink, sms, tps, 8);
        runt<2, 1>("V1 no disc op (7 packed)", CT, sink, sms, tps, 7);
        runt<2, 2>("V2 no FADD2 (7 packed)", CT, sink, sms, tps, 7);
        runt<2, 3>("V3 LOP3 instead of 2 SHF (8 packed)", CT, sink, sms, tps, 8);
        runt<2, 5>("V5 disc as 2 scalar FFMA (7 packed + 2)", CT, sink, sms, tps, 8);
        runt<2, 6>("V6 16 spheres per iteration", CT, sink, sms, tps, 8);
        runt<2, 7>("V7 h chain only (3 packed)", CT, sink, sms, tps, 3);
        runt<2, 8>("V8 6 chain ops + FADD2, xor masks", CT, sink, sms, tps, 7);
        runt<2, 9>("V9 7 packed + FSETP/@P OR", CT, sink, sms, tps, 7);
        runt<4, 9>("V9 7 packed + FSETP/@P OR, 4 rays", CT, sink, sms, tps, 7);
        runt<4, 0>("V0 baseline, 4 rays", CT, sink, sms, tps, 8);
        runt<4, 5>("V5 scalar disc, 4 rays", CT, sink, sms, tps, 8);
        runt<1, 0>("V0 baseline, 1 ray", CT, sink, sms, tps, 8);
    }
    runp<8>("P8  fma(s_i.bcast, UR pair, acc)", in, out, CT, sms);
    runp<11>("P11 fma(s_0.bcast, UR pair, acc)", in, out, CT, sms);
    runp<9>("P9  fadd2(acc, UR pair)", in, out, CT, sms);
    runp<10>("P10 fma(A_i, A_i, acc)", in, out, CT, sms);
    runp<12>("P12 scalar fma(s_i, UR, acc) x2", in, out, CT, sms);
    runp<13>("P13 scalar fma(A_i, A_i, acc) x2", in, out, CT, sms);
    return 0;
}
