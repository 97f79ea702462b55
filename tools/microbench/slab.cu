// slab.cu — throughput of the BVH2 node test (two boxes against one ray) written with different instruction mixes
// (development aid, not part of libptcore).  Each thread keeps a node (8 words) and a ray (slab constants) in registers and
// runs ITER node tests whose results feed a cheap accumulator; 1024 threads per CTA, one CTA per SM.
//   prmt_ffma : the production form — 12 PRMT (16-bit plane -> float) + 12 FFMA + 4 FMNMX + 4 FMNMX3
//   idp_ffma  : the decode as IDP.2A (dp2a: picks a 16-bit half, scales by 128, adds the exponent) — an FMA-pipe instruction
//   prmt_ffma2 / idp_ffma2 : entry and exit plane of an axis in ONE FFMA2 (f32x2)
//   mix_ffma / mix_ffma2 : entry planes decoded by PRMT, exit planes by IDP.2A (the decode split over both pipes)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o slab slab.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITER 2048
__device__ __forceinline__ float u2f(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t sel) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(0x3F000000u), "r"(sel)); return r; }
__device__ __forceinline__ uint32_t idp(uint32_t a, uint32_t sel) { uint32_t r; asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(sel), "r"(0x3F800000u)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

struct Ray { float ax, ay, az, nx, fx, ny, fy, nz, fz; uint32_t sx, sy, sz, tx, ty, tz; };   // s*: entry-plane selector, t*: exit-plane selector

template <int V> __device__ __forceinline__ void box(const Ray& r, uint32_t wx, uint32_t wy, uint32_t wz, float tmax, float& tn, float& tf) {
    float nx, ny, nz, fx, fy, fz;
    if (V == 0 || V == 2) {          // PRMT decode
        uint32_t a = prmt(wx, r.sx), b = prmt(wx, r.tx), c = prmt(wy, r.sy), d = prmt(wy, r.ty), e = prmt(wz, r.sz), f = prmt(wz, r.tz);
        if (V == 0) { nx = fmaf(u2f(a), r.ax, r.nx); fx = fmaf(u2f(b), r.ax, r.fx); ny = fmaf(u2f(c), r.ay, r.ny); fy = fmaf(u2f(d), r.ay, r.fy); nz = fmaf(u2f(e), r.az, r.nz); fz = fmaf(u2f(f), r.az, r.fz); }
        else { float2 x = __ffma2_rn(make_float2(u2f(a), u2f(b)), make_float2(r.ax, r.ax), make_float2(r.nx, r.fx));
               float2 y = __ffma2_rn(make_float2(u2f(c), u2f(d)), make_float2(r.ay, r.ay), make_float2(r.ny, r.fy));
               float2 z = __ffma2_rn(make_float2(u2f(e), u2f(f)), make_float2(r.az, r.az), make_float2(r.nz, r.fz));
               nx = x.x; fx = x.y; ny = y.x; fy = y.y; nz = z.x; fz = z.y; }
    } else if (V == 1 || V == 3) {   // IDP decode
        uint32_t a = idp(wx, r.sx), b = idp(wx, r.tx), c = idp(wy, r.sy), d = idp(wy, r.ty), e = idp(wz, r.sz), f = idp(wz, r.tz);
        if (V == 1) { nx = fmaf(u2f(a), r.ax, r.nx); fx = fmaf(u2f(b), r.ax, r.fx); ny = fmaf(u2f(c), r.ay, r.ny); fy = fmaf(u2f(d), r.ay, r.fy); nz = fmaf(u2f(e), r.az, r.nz); fz = fmaf(u2f(f), r.az, r.fz); }
        else { float2 x = __ffma2_rn(make_float2(u2f(a), u2f(b)), make_float2(r.ax, r.ax), make_float2(r.nx, r.fx));
               float2 y = __ffma2_rn(make_float2(u2f(c), u2f(d)), make_float2(r.ay, r.ay), make_float2(r.ny, r.fy));
               float2 z = __ffma2_rn(make_float2(u2f(e), u2f(f)), make_float2(r.az, r.az), make_float2(r.nz, r.fz));
               nx = x.x; fx = x.y; ny = y.x; fy = y.y; nz = z.x; fz = z.y; }
    } else {                         // entry planes PRMT, exit planes IDP
        uint32_t a = prmt(wx, r.sx), b = idp(wx, r.tx), c = prmt(wy, r.sy), d = idp(wy, r.ty), e = prmt(wz, r.sz), f = idp(wz, r.tz);
        if (V == 4) { nx = fmaf(u2f(a), r.ax, r.nx); fx = fmaf(u2f(b), r.ax, r.fx); ny = fmaf(u2f(c), r.ay, r.ny); fy = fmaf(u2f(d), r.ay, r.fy); nz = fmaf(u2f(e), r.az, r.nz); fz = fmaf(u2f(f), r.az, r.fz); }
        else { float2 x = __ffma2_rn(make_float2(u2f(a), u2f(b)), make_float2(r.ax, r.ax), make_float2(r.nx, r.fx));
               float2 y = __ffma2_rn(make_float2(u2f(c), u2f(d)), make_float2(r.ay, r.ay), make_float2(r.ny, r.fy));
               float2 z = __ffma2_rn(make_float2(u2f(e), u2f(f)), make_float2(r.az, r.az), make_float2(r.nz, r.fz));
               nx = x.x; fx = x.y; ny = y.x; fy = y.y; nz = z.x; fz = z.y; }
    }
    tn = max3(nx, ny, fmaxf(nz, 0.0f));
    tf = min3(fx, fy, fminf(fz, tmax));
}

template <int V> __global__ void __launch_bounds__(1024, 1) k_slab(const uint32_t* __restrict__ nodes, const float* __restrict__ rays, uint32_t* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t w[8];
    for (int i = 0; i < 8; ++i) w[i] = nodes[(t & 1023u) * 8 + i];
    Ray r;
    const float* q = rays + (t & 1023u) * 16;
    r.ax = q[0]; r.ay = q[1]; r.az = q[2]; r.nx = q[3]; r.fx = q[4]; r.ny = q[5]; r.fy = q[6]; r.nz = q[7]; r.fz = q[8];
    r.sx = __float_as_uint(q[9]); r.sy = __float_as_uint(q[10]); r.sz = __float_as_uint(q[11]);
    r.tx = r.sx ^ 0x0220u; r.ty = r.sy ^ 0x0220u; r.tz = r.sz ^ 0x0220u;
    float tmax = q[12];
    uint32_t acc = 0;
    for (int i = 0; i < ITER; ++i) {
        float tn0, tf0, tn1, tf1;
        box<V>(r, w[0], w[1], w[2], tmax, tn0, tf0);
        box<V>(r, w[4], w[5], w[6], tmax, tn1, tf1);
        const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
        const bool take1 = !h0 || (h1 && tn1 < tn0);
        const uint32_t next = take1 ? w[7] : w[3];
        acc += next + (h0 && h1 ? 1u : 0u);
        // perturb the node so the loop cannot be hoisted: rotate words through the chosen child reference (cheap, 3 instructions)
        w[0] ^= next & 0x00010001u; w[4] += (next >> 3) & 0x00010001u; tmax = tmax * 0.999f;
    }
    out[t] = acc;
}

template <int V> static void run(const char* name, const uint32_t* nodes, const float* rays, uint32_t* out, int sms, int khz) {
    k_slab<V><<<sms, 1024>>>(nodes, rays, out);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 3; ++i) k_slab<V><<<sms, 1024>>>(nodes, rays, out);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-12s %8.4f ms   %6.2f cycles per node test per SM sub-partition\n", name, ms, cycles / ((double)ITER * 8));   // 8 warps per scheduler
}

int main() {
    int sms = 0, khz = 0; cudaSetDevice(0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t* nodes; float* rays; uint32_t* out;
    cudaMalloc(&nodes, 1024 * 8 * 4); cudaMalloc(&rays, 1024 * 16 * 4); cudaMalloc(&out, (size_t)sms * 1024 * 4);
    uint32_t hn[1024 * 8]; float hr[1024 * 16];
    uint32_t s = 12345;
    for (int i = 0; i < 1024 * 8; ++i) { s = s * 1664525u + 1013904223u; hn[i] = (s & 0x7FFF7FFFu) | 0x80008000u; }
    for (int i = 0; i < 1024; ++i) { for (int k = 0; k < 9; ++k) { s = s * 1664525u + 1013904223u; hr[i * 16 + k] = (float)(s >> 8) / 16777216.0f - 0.5f; }
        for (int k = 9; k < 12; ++k) { s = s * 1664525u + 1013904223u; uint32_t sel = (s >> 31) ? 0x7104u : 0x7324u; hr[i * 16 + k] = *(float*)&sel; } hr[i * 16 + 12] = 1e9f; }
    cudaMemcpy(nodes, hn, sizeof(hn), cudaMemcpyHostToDevice); cudaMemcpy(rays, hr, sizeof(hr), cudaMemcpyHostToDevice);
    printf("%d SMs, nominal %d MHz, %d node tests per thread, 32 warps per SM\n", sms, khz / 1000, ITER);
    run<0>("prmt_ffma", nodes, rays, out, sms, khz);
    run<1>("idp_ffma", nodes, rays, out, sms, khz);
    run<2>("prmt_ffma2", nodes, rays, out, sms, khz);
    run<3>("idp_ffma2", nodes, rays, out, sms, khz);
    run<4>("mix_ffma", nodes, rays, out, sms, khz);
    run<5>("mix_ffma2", nodes, rays, out, sms, khz);
    return cudaGetLastError() != cudaSuccess;
}
