// Micro-benchmarks behind profiles/README.md: issue rate and latency of the fp32 / packed fp32x2 / select
// instructions the step kernel leans on.  nvcc -arch=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define N_IT 4096
template <int KIND, int ILP>
__global__ void k(float* out, u64* cyc, float a0)
{
    float x[ILP], y[ILP];
    u64 px[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = a0 + i + threadIdx.x; y[i] = 1.0f + i; asm("mov.b64 %0, {%1, %2};" : "=l"(px[i]) : "f"(x[i]), "f"(y[i])); }
    const float c = a0 * 0.5f;
    u64 pc; asm("mov.b64 %0, {%1, %2};" : "=l"(pc) : "f"(c), "f"(c));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < N_IT; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
            if (KIND == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
            if (KIND == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(px[i]) : "l"(pc));
            if (KIND == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(px[i]) : "l"(pc));
            if (KIND == 4) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(c));
            if (KIND == 5) asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %1, %0, p;}" : "+f"(x[i]) : "f"(c));
            if (KIND == 6) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(c));
            if (KIND == 7) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(px[i]) : "l"(pc));
            if (KIND >= 8) {
                int& v = reinterpret_cast<int&>(x[i]);
                const int ci = __float_as_int(c);
                if (KIND == 8) asm volatile("add.s32 %0, %0, %1;" : "+r"(v) : "r"(ci));
                if (KIND == 9) asm volatile("xor.b32 %0, %0, %1;" : "+r"(v) : "r"(ci));
                if (KIND == 10) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(v) : "r"(ci));
                if (KIND == 11) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(v) : "r"(ci));
                if (KIND == 12) asm volatile("{.reg .pred p; setp.lt.s32 p, %0, %1; selp.s32 %0, %1, %0, p;}" : "+r"(v) : "r"(ci));
                if (KIND == 13) asm volatile("min.s32 %0, %0, %1;" : "+r"(v) : "r"(ci));
                if (KIND == 14) asm volatile("popc.b32 %0, %0;" : "+r"(v));
                if (KIND == 15) asm volatile("bfind.u32 %0, %0;" : "+r"(v));
                if (KIND == 16) asm volatile("{.reg .f32 t; rsqrt.approx.f32 t, %0; mov.b32 %0, t;}" : "+f"(x[i]));
                if (KIND == 17) asm volatile("shfl.sync.idx.b32 %0, %0, 3, 31, 0xffffffff;" : "+r"(v));
                if (KIND == 18) asm volatile("{.reg .pred p; .reg .b32 t; setp.ne.s32 p, %0, 0; vote.sync.ballot.b32 t, p, 0xffffffff; or.b32 %0, %0, t;}" : "+r"(v));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(px[i])); s += x[i] + lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int KIND, int ILP>
void run(const char* name, int warps)
{
    float* out; u64* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<KIND, ILP><<<148, warps * 32>>>(out, cyc, 1.0f);
    k<KIND, ILP><<<148, warps * 32>>>(out, cyc, 1.0f);
    cudaDeviceSynchronize();
    u64 h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    // per SM sub-partition: warps/4 warps issue ILP*N_IT instructions each
    printf("%-28s warps/SM %2d ILP %d: %.2f cycles per instruction per warp, %.2f inst/clk/sub-partition\n", name, warps, ILP,
           c / (double)(N_IT * ILP), (warps / 4.0) * N_IT * ILP / c);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<0, 1>("FADD dependent chain", 4);
    run<2, 1>("FADD2 dependent chain", 4);
    run<3, 1>("FMUL2 dependent chain", 4);
    run<4, 1>("FMNMX dependent chain", 4);
    run<5, 1>("FSETP+FSEL dependent chain", 4);
    run<6, 1>("FFMA dependent chain", 4);
    run<0, 8>("FADD throughput", 32);
    run<1, 8>("FMUL throughput", 32);
    run<6, 8>("FFMA throughput", 32);
    run<2, 8>("FADD2 throughput", 32);
    run<3, 8>("FMUL2 throughput", 32);
    run<7, 8>("FFMA2 throughput", 32);
    run<4, 8>("FMNMX throughput", 32);
    run<5, 8>("FSETP+FSEL (2 inst) throughput", 32);
    run<8, 8>("IADD throughput", 32);
    run<9, 8>("LOP3 throughput", 32);
    run<10, 8>("SHF throughput", 32);
    run<11, 8>("IMAD throughput", 32);
    run<12, 8>("ISETP+SEL (2 inst) throughput", 32);
    run<13, 8>("IMNMX throughput", 32);
    run<14, 8>("POPC throughput", 32);
    run<15, 8>("FLO (bfind) throughput", 32);
    run<16, 8>("MUFU.RSQ throughput", 32);
    run<17, 8>("SHFL throughput", 32);
    run<18, 8>("ISETP+VOTE+LOP3 (3 inst) throughput", 32);
    run<8, 1>("IADD dependent chain", 4);
    run<11, 1>("IMAD dependent chain", 4);
    run<12, 1>("ISETP+SEL dependent chain", 4);
    run<14, 1>("POPC dependent chain", 4);
    run<15, 1>("FLO dependent chain", 4);
    run<16, 1>("MUFU.RSQ dependent chain", 4);
    run<17, 1>("SHFL dependent chain", 4);
    run<18, 1>("ISETP+VOTE+LOP3 dependent chain", 4);
    return 0;
}
