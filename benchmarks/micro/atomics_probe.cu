// Micro-benchmark: shared-memory / L2 atomic-min and LDS throughput on the target GPU.
// Design input for the splat kernels (not part of the library).  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/atomics_probe benchmarks/micro/atomics_probe.cu
//   ./gpurun_out/atomics_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int THREADS = 256;
constexpr int ITERS = 256;
constexpr int ZB = 3840;  // u32 slots, like two eyes of a 1920-wide row

// mode 0: lane-consecutive addresses (conflict-free)   mode 1: pseudo-random within ZB
// mode 2: all lanes one address                          mode 3: consecutive + rare collisions (splat-like)
__device__ __forceinline__ int addr_for(int mode, int tid, int it, uint32_t &rng) {
    rng = rng * 1664525u + 1013904223u;
    switch (mode) {
        case 0: return (tid + it * THREADS) % ZB;
        case 1: return (rng >> 8) % ZB;
        case 2: return 7;
        default: return (tid + it * THREADS + ((rng >> 28) == 0 ? 1 : 0)) % ZB;
    }
}

template <int OP>  // 0: ATOMS.MIN u32, 1: atomicMin u64 (shared), 2: plain LDS+STS min (racy), 3: LDS only, 4: STS only
__global__ void smem_probe(int mode, unsigned long long *out_cycles, uint32_t *sink) {
    __shared__ uint32_t zb32[ZB];
    __shared__ unsigned long long zb64[ZB];
    for (int i = threadIdx.x; i < ZB; i += THREADS) { zb32[i] = 0xFFFFFFFFu; zb64[i] = ~0ull; }
    __syncthreads();
    uint32_t rng = threadIdx.x * 2654435761u + blockIdx.x;
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 8
    for (int it = 0; it < ITERS; ++it) {
        const int a = addr_for(mode, threadIdx.x, it, rng);
        const uint32_t key = (rng & 0xFFFF0000u) | threadIdx.x;
        if (OP == 0) atomicMin(&zb32[a], key);
        if (OP == 1) atomicMin(&zb64[a], ((unsigned long long)key << 32) | threadIdx.x);
        if (OP == 2) { if (key < zb32[a]) zb32[a] = key; }
        if (OP == 3) acc += zb32[a];
        if (OP == 4) zb32[a] = key;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc + zb32[threadIdx.x] + (uint32_t)zb64[threadIdx.x];
}

// byte vs word loads from shared memory (the row kernel's de-interleave question)
template <int WIDTH>  // 1: LDS.U8 x3 per pixel, 4: LDS.32, 16: LDS.128
__global__ void lds_probe(unsigned long long *out_cycles, uint32_t *sink) {
    __shared__ __align__(16) uint8_t row[5760 * 2];
    for (int i = threadIdx.x; i < 5760 * 2; i += THREADS) row[i] = (uint8_t)i;
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        if (WIDTH == 1) {
            const int j = (threadIdx.x + it * 7) % 1920;
            acc += row[3 * j] + row[3 * j + 1] + row[3 * j + 2];
        } else if (WIDTH == 4) {
            acc += reinterpret_cast<const uint32_t *>(row)[(threadIdx.x + it * 7) % 1440];
        } else {
            const uint4 v = reinterpret_cast<const uint4 *>(row)[(threadIdx.x * 3 + it) % 360];
            acc += v.x ^ v.y ^ v.z ^ v.w;
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 0x12345678u) sink[0] = acc;
}

// 64-bit RED.MIN into an L2-resident buffer, splat-like pattern: lane-consecutive targets + jitter
__global__ void red64_probe(unsigned long long *zbuf, long long n_slots, int jitter, int iters) {
    uint32_t rng = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
    long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        rng = rng * 1664525u + 1013904223u;
        long long t = (base + (long long)it * gridDim.x * blockDim.x + (jitter ? (rng >> 28) % jitter : 0)) % n_slots;
        atomicMin(&zbuf[t], ((unsigned long long)(rng | 1u) << 32) | (uint32_t)base);
    }
}

template <typename K, typename... A>
static int run_smem(const char *name, K kernel, int ctas_per_sm, int sms, unsigned long long *d_cycles, A... args) {
    const int grid = sms * ctas_per_sm;
    kernel<<<grid, THREADS>>>(args..., d_cycles, (uint32_t *)(d_cycles + 4096));
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    kernel<<<grid, THREADS>>>(args..., d_cycles, (uint32_t *)(d_cycles + 4096));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
    static unsigned long long h[4096];
    CK(cudaMemcpy(h, d_cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
    double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
    // per SM: ctas_per_sm CTAs x 8 warps x ITERS warp-instructions in `mean` cycles (CTAs run concurrently)
    const double warp_instr_per_sm = (double)ctas_per_sm * (THREADS / 32) * ITERS;
    printf("%-44s ctas/sm=%d  cycles=%9.0f  cyc/warp-instr/SM=%7.3f  (%.3f ms)\n", name, ctas_per_sm, mean, mean / warp_instr_per_sm, ms);
    return 0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s  SMs %d  L2 %d MB  clock %d kHz\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, p.clockRate);
    const int sms = p.multiProcessorCount;
    unsigned long long *d_cycles; CK(cudaMalloc(&d_cycles, sizeof(unsigned long long) * 8192));
    const char *modes[] = {"consecutive", "random", "single-address", "splat-like"};
    for (int cps : {1, 4}) {
        for (int m = 0; m < 4; ++m) {
            char nm[96];
            snprintf(nm, sizeof nm, "ATOMS.MIN.u32 %s", modes[m]); if (run_smem(nm, smem_probe<0>, cps, sms, d_cycles, m)) return 1;
            snprintf(nm, sizeof nm, "atomicMin u64 smem %s", modes[m]); if (run_smem(nm, smem_probe<1>, cps, sms, d_cycles, m)) return 1;
            snprintf(nm, sizeof nm, "LDS+cmp+STS (racy) %s", modes[m]); if (run_smem(nm, smem_probe<2>, cps, sms, d_cycles, m)) return 1;
            snprintf(nm, sizeof nm, "LDS.32 gather %s", modes[m]); if (run_smem(nm, smem_probe<3>, cps, sms, d_cycles, m)) return 1;
            snprintf(nm, sizeof nm, "STS.32 scatter %s", modes[m]); if (run_smem(nm, smem_probe<4>, cps, sms, d_cycles, m)) return 1;
        }
        if (run_smem("LDS.U8 x3 / pixel (cyc per 3 loads)", lds_probe<1>, cps, sms, d_cycles)) return 1;
        if (run_smem("LDS.32", lds_probe<4>, cps, sms, d_cycles)) return 1;
        if (run_smem("LDS.128 (48 B lane stride)", lds_probe<16>, cps, sms, d_cycles)) return 1;
    }
    // L2-resident 64-bit RED: 1080p stereo z-buffer = 2 * 1920*1080 slots = 33 MB
    const long long slots = 2ll * 1920 * 1080;
    unsigned long long *zbuf; CK(cudaMalloc(&zbuf, slots * 8));
    CK(cudaMemset(zbuf, 0xFF, slots * 8));
    for (int jitter : {0, 4, 64}) {
        const int grid = sms * 8, iters = 64;
        red64_probe<<<grid, 256>>>(zbuf, slots, jitter, iters);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        red64_probe<<<grid, 256>>>(zbuf, slots, jitter, iters);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        const double n = (double)grid * 256 * iters;
        printf("RED.MIN.u64 L2-resident jitter=%-3d  %.1f M atomics in %.3f ms = %.2f G atomics/s (%.1f GB/s of 8-B keys)\n", jitter,
               n / 1e6, ms, n / ms / 1e6, n * 8 / ms / 1e6);
    }
    return 0;
}
