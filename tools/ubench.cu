// ubench.cu -- sm_100a instruction-throughput probes for the packed-int16 ACS inner loop.
// Not part of the product; it answers "which SASS ops can the butterfly afford" before the
// fused kernel is tuned.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int NCH = 8;   // independent chains per thread

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    if constexpr (OP == 0) return __vadd2(a, b);                       // VIADD.16x2
    if constexpr (OP == 1) return __viaddmin_u16x2(a, b, c);           // VIADDMNMX.U16x2
    if constexpr (OP == 2) return __viaddmin_s16x2(a, b, c);           // VIADDMNMX.S16x2
    if constexpr (OP == 3) return __vmins2(a, b);                 // VIMNMX.S16x2
    if constexpr (OP == 4) return b - a + c;                           // IADD3 with negation
    if constexpr (OP == 5) return (a & b) | c;                         // LOP3
    if constexpr (OP == 6) return __byte_perm(a, b, 0xfdb9);           // PRMT (sign replicate)
    if constexpr (OP == 7) return a * 3u + b;                          // IMAD
    if constexpr (OP == 8) return a + b;                               // IADD3 / IMAD.IADD (compiler's choice)
    if constexpr (OP == 9) return __vmaxu2(a, b);                 // VIMNMX.U16x2
    return a;
}

template <int OP>
__global__ void __launch_bounds__(256) probe(uint32_t *out, uint32_t seed, long long *cyc) {
    uint32_t v[NCH], b = seed * 3 + threadIdx.x, c = seed ^ 0x12345;
#pragma unroll
    for (int i = 0; i < NCH; i++) v[i] = seed + i * 77 + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) v[i] = op<OP>(v[i], b, c);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// The candidate butterfly inner loop: 2 packed butterflies (4 new states) per call.
// A,C: packed old metrics; X,Y: packed branch metrics; K0,K1: packed 0x8000 -/+ delta.
__device__ __forceinline__ void bfly(uint32_t &A, uint32_t &C, uint32_t X, uint32_t Y, uint32_t K0, uint32_t K1,
                                     uint32_t &D0, uint32_t &D1) {
    uint32_t t0 = C + Y;
    uint32_t t1 = C + X;
    D0 = C - A + K0;
    D1 = C - A + K1;
    uint32_t n0 = __viaddmin_u16x2(A, X, t0);
    uint32_t n1 = __viaddmin_u16x2(A, Y, t1);
    A = n0; C = n1;
}

template <int VARIANT>
__global__ void __launch_bounds__(128) probe_bfly(uint32_t *out, uint32_t seed, long long *cyc) {
    // 16 registers of data = 8 butterfly pairs, like one q-column of the fused kernel
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = 0x10001000u + (seed + i * 131 + threadIdx.x) % 977 * 0x10001u;
    uint32_t X = 0x00ff0010u + (seed & 15), Y = 0x01fe01feu - X, K0 = 0x80008000u - (X - Y), K1 = 0x80008000u + (X - Y);
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS / 4; it++) {
#pragma unroll
        for (int st = 0; st < 4; st++) {
            const int sh = 8 >> st;
            uint32_t D[16];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if ((i & sh) == 0) bfly(r[i], r[i | sh], X, Y, K0, K1, D[i], D[i | sh]);
            }
            if (VARIANT >= 1) {
                // gather 32 sign bits: 8 PRMT + 8 LOP3
                uint32_t w = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    uint32_t s = __byte_perm(D[2 * i], D[2 * i + 1], 0xfdb9);
                    w = (~s & (0x01010101u << i)) | w;
                }
                acc ^= w;
            } else {
#pragma unroll
                for (int i = 0; i < 16; i++) acc ^= D[i];
            }
        }
        X ^= 0x00010000u;   // keep the loop from being hoisted
    }
    long long t1 = clock64();
    uint32_t s = acc;
#pragma unroll
    for (int i = 0; i < 16; i++) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void semantics(uint32_t *out) {
    // does VIADDMNMX.S16x2 add in >16 bits (i.e. min(a+b, 32767) saturates) or wrap?
    out[0] = __viaddmin_s16x2(0x7ff07ff0u, 0x00200020u, 0x7fff7fffu);   // 32752+32 -> sat 0x7fff or wrap 0x8010
    out[1] = __viaddmin_u16x2(0xfff0fff0u, 0x00200020u, 0xffffffffu);   // 65520+32 -> 0xffff or wrap 0x0010
    out[2] = __viaddmin_s16x2(0x80108010u, 0xffe0ffe0u, 0x00000000u);   // -32752-32 -> wrap (pos) or sat(neg)
    bool ph, pl;
    out[3] = __vibmin_s16x2(0x00050003u, 0x00050004u, &ph, &pl);        // hi equal, lo a<b
    out[4] = (ph ? 2 : 0) | (pl ? 1 : 0);
    out[5] = __vadd2(0x0001ffffu, 0x00010001u);                         // no carry across halves -> 0x00020000
}

template <int OP> int run(const char *name, uint32_t *out, long long *cyc, int nsm) {
    const int blocks = nsm * 4, threads = 256;   // 1024 threads/SM = 32 warps
    probe<OP><<<blocks, threads>>>(out, 1, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, 2, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < blocks; i++) if (h[i] > mx) mx = h[i];
    double ops_per_sm = 4.0 * threads * ITERS * NCH;
    printf("%-28s %8.1f thread-ops/clk/SM  (cycles %lld, %.3f ms, %.2f GHz eff)\n", name, ops_per_sm / mx, mx, ms, mx / (ms * 1e6));
    return 0;
}

template <int V> int run_bfly(const char *name, uint32_t *out, long long *cyc, int nsm, int bps) {
    const int blocks = nsm * bps, threads = 128;
    probe_bfly<V><<<blocks, threads>>>(out, 1, cyc);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe_bfly<V><<<blocks, threads>>>(out, 2, cyc);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < blocks; i++) if (h[i] > mx) mx = h[i];
    // per iteration-of-4-stages: 16 regs * 2 states * 4 stages = 128 state updates per thread
    double su_per_sm = (double)bps * threads * (ITERS / 4) * 128.0;
    double su_total = su_per_sm * nsm;
    printf("%-28s %8.2f state-updates/clk/SM  (%d blk/SM, cycles %lld, %.3f ms -> %.3e state-updates/s = %.0f kbit/s)\n",
           name, su_per_sm / mx, bps, mx, ms, su_total / (ms * 1e-3), su_total / (ms * 1e-3) / 8388608.0 / 1e3);
    return 0;
}

int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s sm_%d%d, %d SMs, clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    int nsm = p.multiProcessorCount;
    uint32_t *out; long long *cyc;
    CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * 8 * 256)); CK(cudaMalloc(&cyc, sizeof(long long) * 2048));
    semantics<<<1, 1>>>(out); CK(cudaDeviceSynchronize());
    uint32_t h[8]; CK(cudaMemcpy(h, out, 32, cudaMemcpyDeviceToHost));
    printf("semantics: viaddmin_s16x2 ovf=%08x  viaddmin_u16x2 ovf=%08x  s16x2 neg=%08x  vibmin=%08x preds=%u  vadd2=%08x\n",
           h[0], h[1], h[2], h[3], h[4], h[5]);
    run<0>("VIADD.16x2", out, cyc, nsm);
    run<1>("VIADDMNMX.U16x2", out, cyc, nsm);
    run<2>("VIADDMNMX.S16x2", out, cyc, nsm);
    run<3>("VIMNMX.S16x2", out, cyc, nsm);
    run<9>("VIMNMX.U16x2(max)", out, cyc, nsm);
    run<4>("IADD3 (b-a+c)", out, cyc, nsm);
    run<5>("LOP3", out, cyc, nsm);
    run<6>("PRMT", out, cyc, nsm);
    run<7>("IMAD", out, cyc, nsm);
    run<8>("a+b", out, cyc, nsm);
    for (int bps = 2; bps <= 8; bps *= 2) {
        run_bfly<0>("bfly core (no gather)", out, cyc, nsm, bps);
        run_bfly<1>("bfly + PRMT/LOP3 gather", out, cyc, nsm, bps);
    }
    return 0;
}
