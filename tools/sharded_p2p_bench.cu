// sharded_p2p_bench.cu -- best-case communication floor of the STATE-SPACE-SHARDED multi-GPU variant: not part of the product.
//
// With the 2^23 path metrics sharded over G GPUs by their top log2(G) index bits, every fused 8-stage pass ends in a perfect
// shuffle of the whole 16 MiB metric array: each GPU keeps 1/G of its shard and sends (G-1)/G of it to the others, and nobody
// may start the next pass before everything has arrived.  This program times exactly that exchange, in its cheapest possible
// form, WITHOUT any add-compare-select work: one persistent kernel per GPU whose blocks write their shard straight into the
// peers' next-pass buffers (16-byte peer stores over NVLink, contiguous 32 KiB runs -- what an ACS epilogue would emit), then a
// flag barrier (system fence, one flag store per peer, spin on the own flags).  profiles/r01_sharded_exchange_floor.jsonl holds
// the same exchange through NCCL all_to_all.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/sharded_p2p_bench tools/sharded_p2p_bench.cu -lpthread
//   tools/_bin/sharded_p2p_bench G [passes]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <thread>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int MAXG = 8;
constexpr size_t NSTATES = 1u << 23;

struct Peers {
    uint4 *recv[MAXG];              // every GPU's next-pass buffer (its shard: NSTATES / G metrics)
    volatile unsigned *flags[MAXG]; // every GPU's flag array: flags[g][from] = last pass `from` has delivered
};

// every wait gives up after ~2 s and raises *failed: a broken peer mapping must not hang the GPU
constexpr long long SPIN_CLOCKS = 4000000000ll;

__device__ __forceinline__ void grid_barrier(unsigned *counter, unsigned target, int *failed)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const long long t0 = clock64();
        while (*(volatile unsigned *)counter < target && !*(volatile int *)failed)
            if (clock64() - t0 > SPIN_CLOCKS) { *(volatile int *)failed = 1; break; }
        __threadfence();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512) exchange_persist(const uint4 *src, Peers peers, int me, int G, int passes, unsigned *counter, int *failed)
{
    const size_t shard_vec = NSTATES * 2 / 16 / G;          // 16-byte vectors in this GPU's shard
    const size_t chunk_vec = shard_vec / G;                 // ... of which this many go to each GPU
    unsigned bar = 0;
    for (int p = 1; p <= passes && !*(volatile int *)failed; p++) {
        // perfect shuffle as a block transpose: chunk c of my shard becomes chunk `me` of GPU c's next-pass shard
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < shard_vec; i += (size_t)gridDim.x * blockDim.x) {
            const int c = (int)(i / chunk_vec);
            uint4 v = src[i];
            v.x += (unsigned)p;
            peers.recv[c][(size_t)me * chunk_vec + (i - (size_t)c * chunk_vec)] = v;
        }
        __threadfence_system();
        bar += gridDim.x;
        grid_barrier(counter, bar, failed);                  // every block of this GPU has issued (and fenced) its stores
        if (blockIdx.x == 0) {
            if (threadIdx.x < G) peers.flags[threadIdx.x][me] = (unsigned)p;                 // "my data for pass p is with you"
            if (threadIdx.x < G) {                                                           // everybody's data is with me
                const long long t0 = clock64();
                while (peers.flags[me][threadIdx.x] < (unsigned)p && !*(volatile int *)failed)
                    if (clock64() - t0 > SPIN_CLOCKS) { *(volatile int *)failed = 1; break; }
            }
        }
        bar += gridDim.x;
        grid_barrier(counter, bar, failed);
    }
}

int main(int argc, char **argv)
{
    const int G = argc > 1 ? atoi(argv[1]) : 2;
    const int passes = argc > 2 ? atoi(argv[2]) : 4000;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (G < 2 || G > MAXG || G > ndev || (G & (G - 1))) { fprintf(stderr, "need 2, 4 or 8 GPUs (%d visible)\n", ndev); return 1; }
    Peers peers{};
    std::vector<uint4 *> src(G);
    std::vector<unsigned *> counter(G);
    std::vector<int *> failed(G);
    std::vector<int> grid(G);
    const size_t shard_bytes = NSTATES * 2 / G;
    for (int g = 0; g < G; g++) {
        CK(cudaSetDevice(g));
        for (int o = 0; o < G; o++) if (o != g) { cudaError_t e = cudaDeviceEnablePeerAccess(o, 0); if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e); cudaGetLastError(); }
        CK(cudaMalloc(&src[g], shard_bytes));
        CK(cudaMemset(src[g], g + 1, shard_bytes));
        CK(cudaMalloc(&peers.recv[g], shard_bytes));
        unsigned *f;
        CK(cudaMalloc(&f, MAXG * sizeof(unsigned)));
        CK(cudaMemset(f, 0, MAXG * sizeof(unsigned)));
        peers.flags[g] = f;
        CK(cudaMalloc(&counter[g], sizeof(unsigned)));
        CK(cudaMemset(counter[g], 0, sizeof(unsigned)));
        CK(cudaMalloc(&failed[g], sizeof(int)));
        CK(cudaMemset(failed[g], 0, sizeof(int)));
        int sms = 0, per_sm = 0;
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, exchange_persist, 512, 0));
        grid[g] = sms * std::min(per_sm, 2);                // all blocks resident: the grid barrier needs it
        CK(cudaDeviceSynchronize());
    }
    for (int rep = 0; rep < 2; rep++) {
        std::vector<float> ms(G);
        std::vector<std::thread> th;
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g] {
                CK(cudaSetDevice(g));
                CK(cudaMemset(counter[g], 0, sizeof(unsigned)));
                CK(cudaMemset((void *)peers.flags[g], 0, MAXG * sizeof(unsigned)));
                CK(cudaDeviceSynchronize());
            });
        for (auto &t : th) t.join();
        th.clear();
        for (int g = 0; g < G; g++)
            th.emplace_back([&, g] {
                CK(cudaSetDevice(g));
                cudaEvent_t e0, e1;
                CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
                CK(cudaEventRecord(e0));
                exchange_persist<<<grid[g], 512>>>(src[g], peers, g, G, passes, counter[g], failed[g]);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms[g], e0, e1));
            });
        for (auto &t : th) t.join();
        for (int g = 0; g < G; g++) {
            int f = 0;
            CK(cudaSetDevice(g));
            CK(cudaMemcpy(&f, failed[g], sizeof f, cudaMemcpyDeviceToHost));
            if (f) { fprintf(stderr, "GPU %d gave up waiting (peer stores or flags did not arrive)\n", g); return 2; }
        }
        const float worst = *std::max_element(ms.begin(), ms.end());
        const double us = 1e3 * worst / passes;
        const double sent = (double)shard_bytes * (G - 1) / G;
        if (rep == 1)
            printf("{\"G\": %d, \"passes\": %d, \"us_per_pass_exchange_only\": %.2f, \"bytes_sent_per_gpu_per_pass\": %.0f, \"peer_store_GBps_per_gpu\": %.1f, "
                   "\"decoded_bits_per_s_ceiling_whole_job\": %.0f, \"method\": \"persistent kernel per GPU: 16-byte peer stores of the shuffled shard + system fence + "
                   "flag barrier over NVLink, no ACS work; device events, max over GPUs\"}\n",
                   G, passes, us, sent, sent / (us * 1e-6) / 1e9, 8.0 / (us * 1e-6));
    }
    return 0;
}
