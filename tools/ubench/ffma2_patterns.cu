// Micro-benchmark: FFMA2 issue rate vs operand pattern on sm_100a (development aid for DESIGN.md §7).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ffma2_patterns ffma2_patterns.cu
#include <cuda_runtime.h>
#include <cstdio>
#define N_ACC 8
template <int P>
__global__ void __launch_bounds__(128, 5) k(const float* __restrict__ in, int iters, float* out, float4 cst) {
    float2 acc[N_ACC], A[N_ACC], B[N_ACC];
    for (int i = 0; i < N_ACC; ++i) {
        acc[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 1]);
        A[i] = make_float2(in[threadIdx.x + 40 + i], in[threadIdx.x + 41 + i]);
        B[i] = make_float2(in[threadIdx.x + 80 + i], in[threadIdx.x + 81 + i]);
    }
    float s0 = in[threadIdx.x + 200], s1 = in[threadIdx.x + 201];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N_ACC; ++i) {
            if (P == 1) acc[i] = __ffma2_rn(acc[i], A[0], B[0]);                         // 1 varying pair, 2 reused pairs
            if (P == 2) acc[i] = __ffma2_rn(A[i], B[(i + 3) % N_ACC], acc[i]);           // 3 distinct pairs, no reuse
            if (P == 3) acc[i] = __ffma2_rn(A[i], make_float2(s0, s0), acc[i]);          // 2 pairs + scalar bcast (reused)
            if (P == 4) acc[i] = __ffma2_rn(A[i], make_float2(cst.x, cst.x), acc[i]);    // 2 pairs + uniform/const operand
            if (P == 5) acc[i] = __ffma2_rn(A[i], make_float2((i & 1) ? s0 : s1, (i & 1) ? s0 : s1), acc[i]);  // alternating scalars
            if (P == 6) acc[i] = __ffma2_rn(A[i & 1], make_float2(s0, s0), acc[i]);      // A reused every other instr
        }
        if (P == 7) {  // scalar FFMA equivalent of P2: 3 distinct regs, no reuse
#pragma unroll
            for (int i = 0; i < N_ACC; ++i) {
                acc[i].x = fmaf(A[i].x, B[(i + 3) % N_ACC].x, acc[i].x);
                acc[i].y = fmaf(A[i].y, B[(i + 3) % N_ACC].y, acc[i].y);
            }
        }
    }
    float r = 0;
    for (int i = 0; i < N_ACC; ++i) r += acc[i].x + acc[i].y;
    if (r == 1.2345f) out[0] = r;
}
template <int P>
void run(const char* name, const float* in, float* out, int sms) {
    const int iters = 1 << 15, blocks = sms * 5, threads = 128;
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<P><<<blocks, threads>>>(in, iters, out, make_float4(1.0000001f, 0, 0, 0));
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (rep && ms < best) best = ms;
    }
    const double ops = (double)iters * N_ACC * blocks * threads;  // packed instrs (P7: 2 scalar per slot)
    const double tflops = ops * 4 / (best * 1e-3) / 1e12;
    // cycles per warp-instruction per SMSP at 1.965 GHz
    const double warp_instr = ops / 32.0 * (P == 7 ? 2 : 1);
    const double cyc = best * 1e-3 * 1.965e9 * (sms * 4) / warp_instr;
    printf("%-48s %7.2f TFLOP/s  %5.2f cycles per warp-instruction per SMSP\n", name, tflops, cyc);
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *in, *out;
    cudaMalloc(&in, 4096), cudaMalloc(&out, 64);
    cudaMemset(in, 0, 4096);
    run<1>("P1 acc=fma(acc,A0,B0)   1 pair + 2 reused", in, out, p.multiProcessorCount);
    run<2>("P2 acc=fma(Ai,Bj,acc)   3 distinct pairs", in, out, p.multiProcessorCount);
    run<3>("P3 acc=fma(Ai,s.bcast,acc) 2 pairs + scalar", in, out, p.multiProcessorCount);
    run<4>("P4 acc=fma(Ai,UR.bcast,acc) 2 pairs + uniform", in, out, p.multiProcessorCount);
    run<5>("P5 like P3, alternating scalars", in, out, p.multiProcessorCount);
    run<6>("P6 acc=fma(A(i&1),s,acc)", in, out, p.multiProcessorCount);
    run<7>("P7 scalar FFMA, 3 distinct regs", in, out, p.multiProcessorCount);
    return 0;
}
