// Micro-benchmarks that size the integer pipe for the field arithmetic (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I celo_bls_snark_rs_b200/csrc \
//        tools/microbench_fp.cu -o gpurun_out/microbench_fp && gpurun_out/microbench_fp
// Reports: raw IMAD / IMAD.WIDE / IMAD.HI issue rates, Fq377 / Fq761 products per second with the
// modulus as immediates vs. in constant memory, and XYZZ mixed adds per second.
#include <cstdio>
#include <cuda_runtime.h>

#include "ec.cuh"

using namespace b200;

template <int MODE>
__global__ void __launch_bounds__(256) k_imad(uint32_t *out, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t x0 = a, x1 = b, x2 = a ^ b, x3 = a + b, x4 = a * 3, x5 = b * 5, x6 = a * 7, x7 = b * 9;
    uint32_t y0 = 1, y1 = 2, y2 = 3, y3 = 4, y4 = 5, y5 = 6, y6 = 7, y7 = 8;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {   // IMAD lo
#define S(x) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
            S(x0) S(x1) S(x2) S(x3) S(x4) S(x5) S(x6) S(x7)
#undef S
        } else if (MODE == 1) {   // IMAD.HI
#define S(x) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(a), "r"(b));
            S(x0) S(x1) S(x2) S(x3) S(x4) S(x5) S(x6) S(x7)
#undef S
        } else if (MODE == 2) {   // IMAD.WIDE (64-bit accumulate)
#define S(x, y) asm volatile("{ .reg .u64 t; mov.b64 t, {%0, %1}; mad.wide.u32 t, %2, %3, t; mov.b64 {%0, %1}, t; }" : "+r"(x), "+r"(y) : "r"(a), "r"(b));
            S(x0, y0) S(x1, y1) S(x2, y2) S(x3, y3) S(x4, y4) S(x5, y5) S(x6, y6) S(x7, y7)
#undef S
        } else if (MODE == 3) {   // carry-chained wide pairs (what the field product issues)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x0), "+r"(y0) : "r"(a), "r"(b));
#define S(x, y) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x), "+r"(y) : "r"(a), "r"(b));
            S(x1, y1) S(x2, y2) S(x3, y3)
#undef S
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x4), "+r"(y4) : "r"(b), "r"(a));
#define S(x, y) asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x), "+r"(y) : "r"(b), "r"(a));
            S(x5, y5) S(x6, y6) S(x7, y7)
#undef S
        } else {   // IADD3 chain with carries
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(x0) : "r"(a));
#define S(x) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(x) : "r"(b));
            S(x1) S(x2) S(x3) S(x4) S(x5) S(x6) S(x7)
#undef S
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ y0 ^ y1 ^ y2 ^ y3 ^ y4 ^ y5 ^ y6 ^ y7;
}

template <class F>
__global__ void __launch_bounds__(256) k_mul(typename F::Mem *io, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    F x = F::load(io[t]), y = F::load(io[t + 1]);
    for (int i = 0; i < iters; i++) {
        x = x * y;
        y = y * x;
    }
    io[t] = (x + y).store();
}
template <class F>
__global__ void __launch_bounds__(256) k_addsub(typename F::Mem *io, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    F x = F::load(io[t]), y = F::load(io[t + 1]);
    for (int i = 0; i < iters; i++) {
        x = x + y;
        y = y - x;
    }
    io[t] = (x + y).store();
}
template <class F, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_madd(XYZZMem<F> *io, const AffineMem<F> *pts, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::load(io[t]);
    Affine<F> p = Affine<F>::load(pts[t & 1023]);
    for (int i = 0; i < iters; i++) {
        acc.madd(p.x, p.y);
        p.x = p.x + acc.zz;   // keep operands changing
    }
    io[t] = acc.store();
}

template <class K, class... A>
static float time_kernel(K k, dim3 grid, dim3 block, A... args) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<<<grid, block>>>(args...);   // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<<<grid, block>>>(args...);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    return ms;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, %d MHz\n", prop.name, sms, prop.clockRate / 1000);

    void *buf;
    size_t bytes = (size_t)sms * 8 * 256 * 512 + 4096;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0x01, bytes);
    dim3 grid(sms * 8), block(256);
    double threads = (double)sms * 8 * 256;

    const char *names[] = {"IMAD.lo", "IMAD.HI", "IMAD.WIDE", "wide pair chain (cc)", "IADD3.X chain"};
    int iters = 4096;
    float ms;
    ms = time_kernel(k_imad<0>, grid, block, (uint32_t *)buf, 7u, iters);
    printf("%-22s %8.3f ms  %7.2f Tinstr/s (thread-level)\n", names[0], ms, threads * iters * 8 / ms / 1e9);
    ms = time_kernel(k_imad<1>, grid, block, (uint32_t *)buf, 7u, iters);
    printf("%-22s %8.3f ms  %7.2f Tinstr/s\n", names[1], ms, threads * iters * 8 / ms / 1e9);
    ms = time_kernel(k_imad<2>, grid, block, (uint32_t *)buf, 7u, iters);
    printf("%-22s %8.3f ms  %7.2f Tinstr/s\n", names[2], ms, threads * iters * 8 / ms / 1e9);
    ms = time_kernel(k_imad<3>, grid, block, (uint32_t *)buf, 7u, iters);
    printf("%-22s %8.3f ms  %7.2f T wide-mads/s\n", names[3], ms, threads * iters * 8 / ms / 1e9);
    ms = time_kernel(k_imad<4>, grid, block, (uint32_t *)buf, 7u, iters);
    printf("%-22s %8.3f ms  %7.2f Tinstr/s\n", names[4], ms, threads * iters * 8 / ms / 1e9);

    iters = 256;
    ms = time_kernel(k_mul<Fq377>, grid, block, (Fq377::Mem *)buf, iters);
    printf("Fq377 mul (32-bit limbs)  %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);
    ms = time_kernel(k_addsub<Fq377>, grid, block, (Fq377::Mem *)buf, iters * 4);
    printf("Fq377 add+sub             %8.3f ms  %7.2f Gop/s\n", ms, threads * iters * 8 / ms / 1e6);
    ms = time_kernel(k_mul<Fq761>, grid, block, (Fq761::Mem *)buf, iters / 4);
    printf("Fq761 mul (out-of-line)   %8.3f ms  %7.2f Gmul/s\n", ms, threads * (iters / 4) * 2 / ms / 1e6);
    // single-warp latency of a dependent product chain
    ms = time_kernel(k_mul<Fq377>, dim3(1), dim3(32), (Fq377::Mem *)buf, 4096);
    printf("Fq377 mul latency (1 warp, dependent chain) %8.1f ns/mul\n", ms * 1e6 / (4096 * 2));

    iters = 64;
    {
        dim3 g(sms * 12), b(128);
        double th = (double)sms * 12 * 128;
        ms = time_kernel(k_madd<Fq377, 128, 2>, g, b, (XYZZMem<Fq377> *)buf, (const AffineMem<Fq377> *)buf, iters);
        printf("G1-377 madd 128x2      %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 128, 3>, g, b, (XYZZMem<Fq377> *)buf, (const AffineMem<Fq377> *)buf, iters);
        printf("G1-377 madd 128x3      %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 128, 4>, g, b, (XYZZMem<Fq377> *)buf, (const AffineMem<Fq377> *)buf, iters);
        printf("G1-377 madd 128x4      %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 256, 1>, g, b, (XYZZMem<Fq377> *)buf, (const AffineMem<Fq377> *)buf, iters);
        printf("G1-377 madd 256x1      %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
    }
    cudaFree(buf);
    return 0;
}
