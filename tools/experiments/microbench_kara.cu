// Round-2 micro-benchmarks (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I celo_bls_snark_rs_b200/csrc -I tools/experiments \
//        tools/experiments/microbench_kara.cu -o tools/_bin/microbench_kara     (then: gpurun -- tools/_bin/microbench_kara)
// 1. parity of the separated Karatsuba product / lazy two-product reduction (fp_kara.cuh) with the fused product,
// 2. products per second, fused vs Karatsuba, 12 and 24 limbs,
// 3. XYZZ mixed additions per second through one shared out-of-line body, fused vs Karatsuba + lazy Y3,
// 4. FP64-pipe probe: DFMA alone, carry-chained IMAD.WIDE alone, both interleaved in one warp.
#include <cstdio>
#include <cuda_runtime.h>

#include "fp_kara.cuh"
#include "ec.cuh"
#include "smem_fp.cuh"

using namespace b200;

template <class F, int BASE>
__global__ void __launch_bounds__(128) k_check(const typename F::Mem *in, uint32_t n, uint32_t *bad) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    F a = F::load(in[t]), b = F::load(in[(t + 1) % n]), c = F::load(in[(t + 7) % n]), d = F::load(in[(t + 13) % n]);
    a.reduce_once();
    b.reduce_once();
    c.reduce_once();
    d.reduce_once();
    if (t % 5 == 0) b = a;
    if (t % 7 == 0) c = F::zero();
    if (t % 11 == 0) d = F::zero() - F::one();
    F r1 = F::mul_outline(a, b), r2 = kara_mul_outline<BASE>(a, b);
    F d1 = F::mul_outline(a, b) + F::mul_outline(c, d), d2 = kara_dot2_outline<BASE>(a, b, c, d);
    if (!(r1 == r2)) atomicAdd(bad, 1u);
    if (!(d1 == d2)) atomicAdd(bad + 1, 1u);
}

template <class F, int MODE, int BASE>
__global__ void __launch_bounds__(256) k_mul(typename F::Mem *io, int iters) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    F x = F::load(io[t]), y = F::load(io[t + 1]);
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
            x = F::mul_outline(x, y);
            y = F::mul_outline(y, x);
        } else {
            x = kara_mul_outline<BASE>(x, y);
            y = kara_mul_outline<BASE>(y, x);
        }
    }
    io[t] = (x + y).store();
}

// the accumulate kernel's mixed addition (msm.cuh xyzz_madd_shared), exceptional cases included
template <class F, int MODE, int BASE>
__device__ __noinline__ XYZZ<F> madd_body(XYZZ<F> a, F px, F py) {
    if (a.is_inf()) return {px, py, F::one(), F::one()};
    auto mul = [](const F &u, const F &v) { return MODE == 0 ? F::mul_outline(u, v) : kara_mul_outline<BASE>(u, v); };
    F p = mul(px, a.zz) - a.x;
    F r = mul(py, a.zzz) - a.y;
    if (p.is_zero()) return r.is_zero() ? XYZZ<F>::dbl_affine(px, py) : XYZZ<F>::inf();
    F pp = F::sqr_outline(p);
    F ppp = mul(p, pp);
    F q = mul(a.x, pp);
    XYZZ<F> o;
    o.x = F::sqr_outline(r) - ppp - q.dbl();
    if (MODE == 2) o.y = kara_dot2_outline<BASE>(r, q - o.x, a.y.neg(), ppp);
    else o.y = mul(r, q - o.x) - mul(a.y, ppp);
    o.zz = mul(a.zz, pp);
    o.zzz = mul(a.zzz, ppp);
    return o;
}
template <class F, int MODE, int BASE, int THREADS, int MINB, bool CHECK = false>
__global__ void __launch_bounds__(THREADS, MINB) k_madd(XYZZMem<F> *io, const AffineMem<F> *pts, int iters, uint32_t *bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> acc = XYZZ<F>::load(io[t]);
    Affine<F> p = Affine<F>::load(pts[t & 1023]);
    p.x.reduce_once();
    p.y.reduce_once();
    XYZZ<F> ref = acc;
    for (int i = 0; i < iters; i++) {
        acc = madd_body<F, MODE, BASE>(acc, p.x, p.y);
        if (CHECK) ref = madd_body<F, 0, BASE>(ref, p.x, p.y);
        p.x = p.x + acc.zz;   // keep operands changing
    }
    if (CHECK && !(acc.x == ref.x && acc.y == ref.y && acc.zz == ref.zz && acc.zzz == ref.zzz)) atomicAdd(bad, 1u);
    if (!CHECK) io[t] = acc.store();
}

// ---- FP64 probe ----
template <int MODE>
__global__ void __launch_bounds__(256) k_pipes(double *out, uint32_t seed, int iters) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
    uint32_t x0 = a, x1 = b, x2 = a ^ b, x3 = a + b, y0 = 1, y1 = 2, y2 = 3, y3 = 4;
    double f0 = a, f1 = b, f2 = a + 0.5, f3 = b + 0.25, ga = 1.0 + a * 1e-9, gb = 1e-3 * b;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0 || MODE == 2) {   // 4 carry-chained wide multiply-adds
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x0), "+r"(y0) : "r"(a), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x1), "+r"(y1) : "r"(a), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x2), "+r"(y2) : "r"(a), "r"(b));
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x3), "+r"(y3) : "r"(a), "r"(b));
        }
        if (MODE == 1 || MODE == 2) {   // 4 independent DFMAs
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f0) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f1) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f2) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f3) : "d"(ga), "d"(gb));
        }
        if (MODE == 3) {                // 8 DFMAs
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f0) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f1) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f2) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f3) : "d"(ga), "d"(gb));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f0) : "d"(gb), "d"(ga));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f1) : "d"(gb), "d"(ga));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f2) : "d"(gb), "d"(ga));
            asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(f3) : "d"(gb), "d"(ga));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = f0 + f1 + f2 + f3 + (double)(x0 ^ x1 ^ x2 ^ x3 ^ y0 ^ y1 ^ y2 ^ y3);
}


// the shared-memory-resident form of the same mixed addition (smem_fp.cuh)
template <class F, int THREADS, int MINB, bool KARA, bool CHECK>
__global__ void __launch_bounds__(THREADS, MINB) k_madd_slots(XYZZMem<F> *io, const AffineMem<F> *pts, int iters, uint32_t *bad) {
    extern __shared__ uint4 sm[];
    using S = Slots<F, THREADS>;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * 16u;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    XYZZ<F> ref = XYZZ<F>::load(io[t]);
    S::st(base, S::X, ref.x);
    S::st(base, S::Y, ref.y);
    S::st(base, S::ZZ, ref.zz);
    S::st(base, S::ZZZ, ref.zzz);
    bool inf = false;
    uint32_t idx = (uint32_t)t * 7u;
    S::fetch(base, S::PX, &pts[idx & 1023].x);
    S::fetch(base, S::PY, &pts[idx & 1023].y);
    cp_async_commit();
    for (int i = 0; i < iters; i++) {
        cp_async_wait_all();
        if (CHECK) ref = madd_body<F, 0, 6>(ref, S::ld(base, S::PX), S::ld(base, S::PY));
        slot_madd<F, THREADS, KARA>(base, inf, [&]() {
            idx += 13u;
            S::fetch(base, S::PX, &pts[idx & 1023].x);
            S::fetch(base, S::PY, &pts[idx & 1023].y);
            cp_async_commit();
        });
    }
    cp_async_wait_all();
    XYZZ<F> acc = {S::ld(base, S::X), S::ld(base, S::Y), S::ld(base, S::ZZ), S::ld(base, S::ZZZ)};
    if (CHECK && (inf || !(acc.x == ref.x && acc.y == ref.y && acc.zz == ref.zz && acc.zzz == ref.zzz))) atomicAdd(bad, 1u);
    if (!CHECK) io[t] = acc.store();
}
template <class F, int THREADS, int MINB, bool KARA>
static void run_slots(const char *name, void *buf, uint32_t *bad, int sms, int iters) {
    using S = Slots<F, THREADS>;
    const AffineMem<F> *pts = (const AffineMem<F> *)((char *)buf + (64 << 20));
    auto kc = k_madd_slots<F, THREADS, MINB, KARA, true>;
    auto kt = k_madd_slots<F, THREADS, MINB, KARA, false>;
    cudaFuncSetAttribute(kc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES);
    cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::BYTES);
    cudaMemset(bad, 0, 16);
    kc<<<64, THREADS, S::BYTES>>>((XYZZMem<F> *)buf, pts, 8, bad);
    uint32_t h = 0;
    cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost);
    int blocks = sms * MINB * 4;
    double th = (double)blocks * THREADS;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kt<<<blocks, THREADS, S::BYTES>>>((XYZZMem<F> *)buf, pts, iters, (uint32_t *)nullptr);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kt<<<blocks, THREADS, S::BYTES>>>((XYZZMem<F> *)buf, pts, iters, (uint32_t *)nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("%-34s %8.3f ms  %7.3f Gmadd/s  (parity mismatches %u of %d, smem %u B/block) %s\n", name, ms, th * iters / ms / 1e6, h,
           64 * THREADS, (unsigned)S::BYTES, err == cudaSuccess ? "" : cudaGetErrorString(err));
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

template <class F, int BASE>
static void run_check(const char *name, void *buf, uint32_t *bad) {
    cudaMemset(bad, 0, 16);
    const uint32_t n = 1u << 18;
    k_check<F, BASE><<<n / 128, 128>>>((const typename F::Mem *)buf, n, bad);
    k_madd<F, 2, BASE, 128, 1, true><<<256, 128>>>((XYZZMem<F> *)buf, (const AffineMem<F> *)((char *)buf + (64 << 20)), 6, bad + 2);
    uint32_t h[4];
    cudaMemcpy(h, bad, 16, cudaMemcpyDeviceToHost);
    cudaError_t err = cudaGetLastError();
    printf("parity %-22s mul mismatches %u, dot2 mismatches %u, madd(kara+lazy) mismatches %u of %u / %u / 32768 %s\n", name, h[0],
           h[1], h[2], n, n, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, %d MHz\n", prop.name, sms, prop.clockRate / 1000);

    void *buf;
    size_t bytes = (size_t)sms * 12 * 256 * 512 + (128 << 20);
    cudaMalloc(&buf, bytes);
    // pseudo-random words (top limbs are cut below p by reduce_once only approximately: the kernels reduce twice... inputs are
    // arbitrary 32-bit patterns; the product bounds need < 2^(32 N - 7), so clear the top byte of every word group below)
    {
        size_t words = bytes / 4;
        uint32_t *h = (uint32_t *)malloc(bytes);
        uint64_t s = 0x9e3779b97f4a7c15ull;
        for (size_t i = 0; i < words; i++) {
            s ^= s << 13;
            s ^= s >> 7;
            s ^= s << 17;
            h[i] = (uint32_t)(s >> 16);
            if (i % 12 == 11) h[i] &= 0x00ffffffu;   // below 2^376 (12-limb view) and below 2^760 (24-limb view: words 23, 47, ...)
        }
        cudaMemcpy(buf, h, bytes, cudaMemcpyHostToDevice);
        free(h);
    }
    uint32_t *bad;
    cudaMalloc(&bad, 16);
    run_check<Fq377, 6>("Fq377 BASE 6", buf, bad);
    run_check<Fq377, 12>("Fq377 BASE 12 (no K)", buf, bad);
    run_check<Fq761, 12>("Fq761 BASE 12", buf, bad);
    run_check<Fq761, 6>("Fq761 BASE 6", buf, bad);

    dim3 grid(sms * 8), block(256);
    double threads = (double)sms * 8 * 256;
    float ms;
    int iters = 256;
    ms = time_kernel(k_mul<Fq377, 0, 6>, grid, block, (Fq377::Mem *)buf, iters);
    printf("Fq377 mul fused (out of line)      %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);
    ms = time_kernel(k_mul<Fq377, 1, 6>, grid, block, (Fq377::Mem *)buf, iters);
    printf("Fq377 mul Karatsuba 6 + redc       %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);
    ms = time_kernel(k_mul<Fq377, 1, 12>, grid, block, (Fq377::Mem *)buf, iters);
    printf("Fq377 mul schoolbook 12 + redc     %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);
    iters = 64;
    ms = time_kernel(k_mul<Fq761, 0, 6>, grid, block, (Fq761::Mem *)buf, iters);
    printf("Fq761 mul fused (out of line)      %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);
    ms = time_kernel(k_mul<Fq761, 1, 12>, grid, block, (Fq761::Mem *)buf, iters);
    printf("Fq761 mul Karatsuba 12 + redc      %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);
    ms = time_kernel(k_mul<Fq761, 1, 6>, grid, block, (Fq761::Mem *)buf, iters);
    printf("Fq761 mul Karatsuba 12/6 + redc    %8.3f ms  %7.2f Gmul/s\n", ms, threads * iters * 2 / ms / 1e6);

    iters = 64;
    {
        dim3 g(sms * 12), b(128);
        double th = (double)sms * 12 * 128;
        const AffineMem<Fq377> *pts = (const AffineMem<Fq377> *)((char *)buf + (64 << 20));
        ms = time_kernel(k_madd<Fq377, 0, 6, 128, 3>, g, b, (XYZZMem<Fq377> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-377 madd fused 128x3            %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 1, 6, 128, 3>, g, b, (XYZZMem<Fq377> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-377 madd Karatsuba 128x3        %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 2, 6, 128, 3>, g, b, (XYZZMem<Fq377> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-377 madd Karatsuba+lazy 128x3   %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 2, 6, 128, 2>, g, b, (XYZZMem<Fq377> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-377 madd Karatsuba+lazy 128x2   %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq377, 2, 12, 128, 3>, g, b, (XYZZMem<Fq377> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-377 madd schoolbook+lazy 128x3  %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
    }
    iters = 16;
    {
        dim3 g(sms * 8), b(128);
        double th = (double)sms * 8 * 128;
        const AffineMem<Fq761> *pts = (const AffineMem<Fq761> *)((char *)buf + (64 << 20));
        ms = time_kernel(k_madd<Fq761, 0, 6, 128, 2>, g, b, (XYZZMem<Fq761> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-761 madd fused 128x2            %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq761, 2, 12, 128, 2>, g, b, (XYZZMem<Fq761> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-761 madd Karatsuba 12 + lazy    %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq761, 2, 6, 128, 2>, g, b, (XYZZMem<Fq761> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-761 madd Karatsuba 12/6 + lazy  %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
        ms = time_kernel(k_madd<Fq761, 1, 6, 128, 2>, g, b, (XYZZMem<Fq761> *)buf, pts, iters, (uint32_t *)nullptr);
        printf("G1-761 madd Karatsuba 12/6         %8.3f ms  %7.2f Gmadd/s\n", ms, th * iters / ms / 1e6);
    }
    run_slots<Fq377, 128, 3, false>("G1-377 slots fused 128x3", buf, bad, sms, 64);
    run_slots<Fq377, 128, 3, true>("G1-377 slots Karatsuba 128x3", buf, bad, sms, 64);
    run_slots<Fq377, 384, 1, false>("G1-377 slots fused 384x1", buf, bad, sms, 64);
    run_slots<Fq377, 256, 1, false>("G1-377 slots fused 256x1", buf, bad, sms, 64);
    run_slots<Fq761, 224, 1, false>("G1-761 slots fused 224x1", buf, bad, sms, 16);
    run_slots<Fq761, 224, 1, true>("G1-761 slots Karatsuba 224x1", buf, bad, sms, 16);
    run_slots<Fq761, 128, 1, false>("G1-761 slots fused 128x1", buf, bad, sms, 16);
    {
        iters = 4096;
        const char *names[] = {"4 x IMAD.WIDE.X chain", "4 x DFMA", "4 x IMAD.WIDE.X + 4 x DFMA", "8 x DFMA"};
        float t[4];
        t[0] = time_kernel(k_pipes<0>, grid, block, (double *)buf, 7u, iters);
        t[1] = time_kernel(k_pipes<1>, grid, block, (double *)buf, 7u, iters);
        t[2] = time_kernel(k_pipes<2>, grid, block, (double *)buf, 7u, iters);
        t[3] = time_kernel(k_pipes<3>, grid, block, (double *)buf, 7u, iters);
        for (int i = 0; i < 4; i++) printf("%-28s %8.3f ms  (%.2f T thread-instr/s)\n", names[i], t[i], threads * iters * (i >= 2 ? 8 : 4) / t[i] / 1e9);
        printf("co-issue: IMAD alone %.3f + DFMA alone %.3f = %.3f ms serial; measured together %.3f ms\n", t[0], t[1], t[0] + t[1], t[2]);
    }
    cudaFree(buf);
    return 0;
}
