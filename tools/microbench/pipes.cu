// pipes.cu — issue-rate microbenchmark of the instructions the BVH slab test is made of (development aid, not part of libptcore).
// Every kernel runs ITER x 8 independent chains per thread, 1024 threads per CTA (32 warps per SM: 8 per scheduler), one CTA per SM,
// and reports warp-instructions per cycle per SM sub-partition.  Pairs of instruction kinds tell whether they share a pipe:
// if a + b together run as fast as the slower of the two alone, they do not.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

#define ITER 4096

#define OP_PRMT(x)   asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(p1), "r"(p2));
#define OP_IDP(x)    asm volatile("dp2a.lo.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(p1), "r"(p2));
#define OP_IDPH(x)   asm volatile("dp2a.hi.u32.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(p1), "r"(p2));
#define OP_FFMA(x)   asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float*)&x) : "f"(f1), "f"(f2));
#define OP_FMNMX(x)  asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float*)&x) : "f"(f1));
#define OP_FMNMX3(x) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(*(float*)&x) : "f"(f1), "f"(f2));
#define OP_LOP(x)    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(p1), "r"(p2));
#define OP_IMAD(x)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(p1), "r"(p2));
#define OP_SHL(x)    asm volatile("shl.b32 %0, %0, 3;" : "+r"(x));
#define OP_FSETP(x)  asm volatile("{ .reg .pred q; setp.lt.f32 q, %0, %1; selp.b32 %0, %0, %2, q; }" : "+r"(x) : "f"(f1), "r"(p1));
#define OP_I2F(x)    asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(*(float*)&x) : "r"(x));
#define OP_VIMNMX(x) asm volatile("min.u32 %0, %0, %1;" : "+r"(x) : "r"(p1));

#define KERNEL(NAME, BODY)                                                                   \
__global__ void __launch_bounds__(1024, 1) NAME(uint32_t* out, uint32_t p1, uint32_t p2, float f1, float f2, int pp_) { \
    uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7; \
    const int pp = pp_; (void)pp;                                                                     \
    for (int i = 0; i < ITER; ++i) { BODY }                                                   \
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;     \
}
#define ALL8(OP) OP(a0) OP(a1) OP(a2) OP(a3) OP(a4) OP(a5) OP(a6) OP(a7)
#define MIX44(OPA, OPB) OPA(a0) OPB(a1) OPA(a2) OPB(a3) OPA(a4) OPB(a5) OPA(a6) OPB(a7)

// FFMA2 needs a register pair
__global__ void __launch_bounds__(1024, 1) k_ffma2(uint32_t* out, uint32_t p1, uint32_t p2, float f1, float f2, int) {
    float2 a0 = make_float2(threadIdx.x, 1.f), a1 = make_float2(2.f, threadIdx.x), a2 = a0, a3 = a1, a4 = a0, a5 = a1, a6 = a0, a7 = a1;
    const float2 c = make_float2(f2, f1);
    for (int i = 0; i < ITER; ++i) {
        a0 = __ffma2_rn(a0, make_float2(f1, f1), c); a1 = __ffma2_rn(a1, make_float2(f1, f1), c); a2 = __ffma2_rn(a2, make_float2(f1, f1), c); a3 = __ffma2_rn(a3, make_float2(f1, f1), c);
        a4 = __ffma2_rn(a4, make_float2(f1, f1), c); a5 = __ffma2_rn(a5, make_float2(f1, f1), c); a6 = __ffma2_rn(a6, make_float2(f1, f1), c); a7 = __ffma2_rn(a7, make_float2(f1, f1), c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + a4.x + a4.y + a5.x + a5.y + a6.x + a6.y + a7.x + a7.y);
}
// FFMA2 + PRMT (4 + 4)
__global__ void __launch_bounds__(1024, 1) k_ffma2_prmt(uint32_t* out, uint32_t p1, uint32_t p2, float f1, float f2, int) {
    float2 a0 = make_float2(threadIdx.x, 1.f), a2 = a0, a4 = a0, a6 = a0;
    uint32_t a1 = threadIdx.x, a3 = a1 + 1, a5 = a1 + 2, a7 = a1 + 3;
    const float2 c = make_float2(f2, f1);
    for (int i = 0; i < ITER; ++i) {
        a0 = __ffma2_rn(a0, make_float2(f1, f1), c); OP_PRMT(a1) a2 = __ffma2_rn(a2, make_float2(f1, f1), c); OP_PRMT(a3)
        a4 = __ffma2_rn(a4, make_float2(f1, f1), c); OP_PRMT(a5) a6 = __ffma2_rn(a6, make_float2(f1, f1), c); OP_PRMT(a7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(a0.x + a0.y + a2.x + a2.y + a4.x + a4.y + a6.x + a6.y) ^ a1 ^ a3 ^ a5 ^ a7;
}
// FFMA2 + FFMA (4 + 4): do they share the pipe?
__global__ void __launch_bounds__(1024, 1) k_ffma2_ffma(uint32_t* out, uint32_t p1, uint32_t p2, float f1, float f2, int) {
    float2 a0 = make_float2(threadIdx.x, 1.f), a2 = a0, a4 = a0, a6 = a0;
    uint32_t a1 = threadIdx.x, a3 = a1 + 1, a5 = a1 + 2, a7 = a1 + 3;
    const float2 c = make_float2(f2, f1);
    for (int i = 0; i < ITER; ++i) {
        a0 = __ffma2_rn(a0, make_float2(f1, f1), c); OP_FFMA(a1) a2 = __ffma2_rn(a2, make_float2(f1, f1), c); OP_FFMA(a3)
        a4 = __ffma2_rn(a4, make_float2(f1, f1), c); OP_FFMA(a5) a6 = __ffma2_rn(a6, make_float2(f1, f1), c); OP_FFMA(a7)
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = __float_as_uint(a0.x + a0.y + a2.x + a2.y + a4.x + a4.y + a6.x + a6.y) ^ a1 ^ a3 ^ a5 ^ a7;
}

KERNEL(k_prmt, ALL8(OP_PRMT))
KERNEL(k_idp, ALL8(OP_IDP))
KERNEL(k_ffma, ALL8(OP_FFMA))
KERNEL(k_fmnmx, ALL8(OP_FMNMX))
KERNEL(k_fmnmx3, ALL8(OP_FMNMX3))
KERNEL(k_lop, ALL8(OP_LOP))
KERNEL(k_imad, ALL8(OP_IMAD))
KERNEL(k_shl, ALL8(OP_SHL))
KERNEL(k_fsetp_sel, ALL8(OP_FSETP))
KERNEL(k_i2f, ALL8(OP_I2F))
KERNEL(k_vimnmx, ALL8(OP_VIMNMX))
KERNEL(k_prmt_ffma, MIX44(OP_PRMT, OP_FFMA))
KERNEL(k_prmt_idp, MIX44(OP_PRMT, OP_IDP))
KERNEL(k_idp_ffma, MIX44(OP_IDP, OP_FFMA))
KERNEL(k_idp_fmnmx, MIX44(OP_IDP, OP_FMNMX))
KERNEL(k_prmt_fmnmx, MIX44(OP_PRMT, OP_FMNMX))
KERNEL(k_fmnmx_ffma, MIX44(OP_FMNMX, OP_FFMA))
KERNEL(k_imad_ffma, MIX44(OP_IMAD, OP_FFMA))
KERNEL(k_imad_prmt, MIX44(OP_IMAD, OP_PRMT))
KERNEL(k_i2f_prmt, MIX44(OP_I2F, OP_PRMT))

typedef void (*kfn)(uint32_t*, uint32_t, uint32_t, float, float, int);
struct Entry { const char* name; kfn fn; int ops; };

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaSetDevice(dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    uint32_t* out; cudaMalloc(&out, (size_t)sms * 1024 * 4);
    Entry es[] = {
        {"PRMT", k_prmt, 8}, {"IDP.2A", k_idp, 8}, {"FFMA", k_ffma, 8}, {"FFMA2", k_ffma2, 8}, {"FMNMX", k_fmnmx, 8}, {"FMNMX3", k_fmnmx3, 8},
        {"LOP3", k_lop, 8}, {"IMAD", k_imad, 8}, {"SHL", k_shl, 8}, {"FSETP+SEL", k_fsetp_sel, 16}, {"I2F.U32", k_i2f, 8}, {"VIMNMX", k_vimnmx, 8},
        {"PRMT+FFMA", k_prmt_ffma, 8}, {"PRMT+IDP.2A", k_prmt_idp, 8}, {"IDP.2A+FFMA", k_idp_ffma, 8}, {"IDP.2A+FMNMX", k_idp_fmnmx, 8},
        {"PRMT+FMNMX", k_prmt_fmnmx, 8}, {"FMNMX+FFMA", k_fmnmx_ffma, 8}, {"IMAD+FFMA", k_imad_ffma, 8}, {"IMAD+PRMT", k_imad_prmt, 8},
        {"I2F+PRMT", k_i2f_prmt, 8}, {"FFMA2+PRMT", k_ffma2_prmt, 8}, {"FFMA2+FFMA", k_ffma2_ffma, 8},
    };
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%d SMs, nominal %d MHz; 32 warps per SM (8 per scheduler), %d x 8 independent chains per thread\n", sms, khz / 1000, ITER);
    printf("%-16s %10s %14s\n", "instruction(s)", "ms", "warp-inst/clk/SMSP (at the nominal clock)");
    for (auto& e : es) {
        e.fn<<<sms, 1024>>>(out, 0x3210, 0x7654, 1.0001f, 0.5f, 1);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        for (int r = 0; r < 3; ++r) e.fn<<<sms, 1024>>>(out, 0x3210, 0x7654, 1.0001f, 0.5f, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
        const double warp_inst = (double)ITER * e.ops * 32;            // per SM: 32 warps
        const double cycles = ms * 1e-3 * khz * 1e3;
        printf("%-16s %10.4f %14.3f\n", e.name, ms, warp_inst / cycles / 4.0);
    }
    if (cudaGetLastError() != cudaSuccess) { printf("CUDA error\n"); return 1; }
    return 0;
}
