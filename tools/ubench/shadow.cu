// Micro-benchmark: which instruction classes issue "in the shadow" of a packed FFMA2 on sm_100a?
// Loop body = 32 FFMA2 (scalar x UR pair + pair, like the sweep) interleaved with NX instructions of
// class X; reports the marginal cost of one X in cycles per warp per SMSP.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o shadow shadow.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
struct Cst { float4 g[64]; };
enum { X_NONE, X_SHF, X_LOP3, X_IADD3, X_IMAD, X_FADD, X_FMUL, X_MUFU, X_LDS, X_PRMT, X_FMNMX, X_FSETP_SEL, X_MOV, X_LEA, X_I2F, X_HFMA2, X_POPC, X_VOTE, X_DFMA, X_DADD, X_DFMA_UR, X_DFMA_CHAIN };

template <int X>
__device__ __forceinline__ void xop(unsigned& a, unsigned b, float& fa, float fb, const float* sm, double& da, double db, double ucst = 1.0) {
    if (X == X_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(da) : "d"(db));
    if (X == X_DFMA_UR) da = fma(da, ucst, ucst);
    if (X == X_DFMA_CHAIN) da = fma(da, db, da * 0.5);
    if (X == X_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(da) : "d"(db));
    if (X == X_SHF) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a) : "r"(b));
    if (X == X_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(b + 1));
    if (X == X_IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (X == X_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a) : "r"(b));
    if (X == X_FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(fa) : "f"(fb));
    if (X == X_FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(fa) : "f"(fb));
    if (X == X_MUFU) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(fa));
    if (X == X_LDS) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(fa) : "r"((unsigned)__cvta_generic_to_shared(sm) + (a & 124u)));
    if (X == X_PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7531;" : "+r"(a) : "r"(b));
    if (X == X_FMNMX) asm volatile("max.f32 %0, %0, %1;" : "+f"(fa) : "f"(fb));
    if (X == X_FSETP_SEL) asm volatile("{.reg .pred p; setp.ge.f32 p, %1, 0f00000000; @p or.b32 %0, %0, 4;}" : "+r"(a) : "f"(fb));
    if (X == X_MOV) asm volatile("mov.b32 %0, %1;" : "=r"(a) : "r"(b));
    if (X == X_LEA) asm volatile("mad.lo.u32 %0, %0, 2, %1;" : "+r"(a) : "r"(b));
    if (X == X_I2F) asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(fa) : "r"(a));
    if (X == X_HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(a) : "r"(b));
    if (X == X_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(a));
    if (X == X_VOTE) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "+r"(a));
}

template <int X, int NX>
__global__ void __launch_bounds__(128, 6) k(const float* __restrict__ in, int iters, float* out,
                                            const __grid_constant__ Cst C) {
    __shared__ float sm[128];
    sm[threadIdx.x] = in[threadIdx.x];
    float2 acc[16];
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 1]);
    float s[4] = {in[threadIdx.x + 200], in[threadIdx.x + 201], in[threadIdx.x + 202], in[threadIdx.x + 203]};
    unsigned xa[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3};
    const unsigned xb[4] = {threadIdx.x * 3 + 1, threadIdx.x * 5 + 1, threadIdx.x * 7 + 1, threadIdx.x * 9 + 1};
    float xf[4] = {s[0], s[1], s[2], s[3]};
    double xd[8] = {s[0], s[1], s[2], s[3], s[0] + 1., s[1] + 1., s[2] + 1., s[3] + 1.};
    const double db = 1.0 + s[0];
    const int b = (iters & 31);
    const float4 g0 = C.g[b], g1 = C.g[b + 1];
    const float2 u[4] = {make_float2(g0.x, g0.y), make_float2(g0.z, g0.w), make_float2(g1.x, g1.y), make_float2(g1.z, g1.w)};
    const double ucd = (double)g0.x * 1.0000001;
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            acc[i & 15] = __ffma2_rn(make_float2(s[i & 3], s[i & 3]), u[i & 3], acc[i & 15]);
            if (NX > 0 && (i % (32 / (NX > 32 ? 32 : NX))) == 0) {
#pragma unroll
                for (int r = 0; r < (NX > 32 ? NX / 32 : 1); ++r) xop<X>(xa[(i + r) & 3], xb[(i + r) & 3], xf[(i + r) & 3], s[(i + r + 1) & 3], sm, xd[(i + r) & 7], db, ucd);
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += acc[i].x + acc[i].y;
    for (int i = 0; i < 4; ++i) r += xf[i] + (float)xa[i];
    for (int i = 0; i < 8; ++i) r += (float)xd[i];
    if (r == 1.2345f) out[0] = r;
}
template <int X, int NX>
double run(const float* in, float* out, const Cst& C, int sms) {
    const int iters = 1 << 13, blocks = sms * 6, threads = 128;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        k<X, NX><<<blocks, threads>>>(in, iters, out, C);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    return best * 1e-3 * 1.965e9 * (sms * 4) / ((double)iters * blocks * threads / 32.0);
}
template <int X>
void sweep(const char* name, const float* in, float* out, const Cst& C, int sms, double base) {
    const double c8 = run<X, 8>(in, out, C, sms), c16 = run<X, 16>(in, out, C, sms), c32 = run<X, 32>(in, out, C, sms),
                 c64 = run<X, 64>(in, out, C, sms);
    printf("%-12s +8: %6.2f (%5.2f/op)  +16: %6.2f (%5.2f/op)  +32: %6.2f (%5.2f/op)  +64: %6.2f (%5.2f/op)\n", name, c8,
           (c8 - base) / 8, c16, (c16 - base) / 16, c32, (c32 - base) / 32, c64, (c64 - base) / 64);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *in, *out;
    cudaMalloc(&in, 4096), cudaMalloc(&out, 64);
    cudaMemset(in, 0, 4096);
    static Cst C;
    for (int i = 0; i < 64; ++i) C.g[i] = make_float4(1, 1, 1, 1);
    const int sms = p.multiProcessorCount;
    const double base = run<X_NONE, 0>(in, out, C, sms);
    printf("base: 32 FFMA2 per iteration = %.2f cycles (%.3f per FFMA2)\n", base, base / 32);
    sweep<X_SHF>("SHF", in, out, C, sms, base);
    sweep<X_IADD3>("IADD3", in, out, C, sms, base);
    sweep<X_FMNMX>("FMNMX", in, out, C, sms, base);
    sweep<X_DFMA>("DFMA", in, out, C, sms, base);
    sweep<X_DADD>("DADD", in, out, C, sms, base);
    sweep<X_DFMA_UR>("DFMA UR", in, out, C, sms, base);
    return 0;
}
