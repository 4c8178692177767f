// Micro-benchmark behind k_scale_count's ranking step: cost of MATCH.ANY against ballots and a shuffle network,
// at the occupancy of that kernel (24 warps per SM).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o
// match_probe match_probe.cu ; prints SM cycles per warp-level operation (throughput view) and per dependent op.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kIters = 4096;

__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int MODE>
__global__ void probe(uint32_t *out, int distinct_mask, long long *cycles) {
    uint32_t x = mix(threadIdx.x + blockIdx.x * 977u);
    uint32_t acc = 0;
    const int lane = threadIdx.x & 31;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < kIters; ++it) {
        x = x * 1664525u + 1013904223u;
        const uint32_t key = (x >> 12) & (uint32_t)distinct_mask;          // ~distinct_mask+1 possible cells
        if (MODE == 0) {                      // MATCH.ANY
            acc += __popc(__match_any_sync(0xffffffffu, key) & ((1u << lane) - 1u));
        } else if (MODE == 1) {               // 12 ballots
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 12; ++b) {
                const uint32_t v = __ballot_sync(0xffffffffu, (key >> b) & 1u);
                m &= ((key >> b) & 1u) ? v : ~v;
            }
            acc += __popc(m & ((1u << lane) - 1u));
        } else if (MODE == 2) {               // bitonic sort of (key << 5 | lane) across the warp: 15 compare-exchange steps
            uint32_t k = (key << 5) | (uint32_t)lane;
#pragma unroll
            for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, k, stride);
                    const bool up = ((lane & size) == 0) || size == 32;
                    const bool lower = (lane & stride) == 0;
                    k = (lower == up) ? min(k, o) : max(k, o);
                }
            }
            const uint32_t prev = __shfl_up_sync(0xffffffffu, k, 1);
            const uint32_t heads = __ballot_sync(0xffffffffu, lane == 0 || (prev >> 5) != (k >> 5));
            acc += lane - (31 - __clz(heads & ((2u << lane) - 1u)));
        } else if (MODE == 3) {               // one shuffle (throughput reference)
            acc += __shfl_xor_sync(0xffffffffu, key, 1);
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int blocks_per_sm, int mask) {
    int dev = 0, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * blocks_per_sm, threads = 128;
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4); cudaMalloc(&cyc, blocks * sizeof(long long));
    probe<MODE><<<blocks, threads>>>(out, mask, cyc);
    probe<MODE><<<blocks, threads>>>(out, mask, cyc);
    cudaDeviceSynchronize();
    long long *h = new long long[blocks];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    const double per_warp_op = avg / kIters;                              // cycles a warp needs per operation
    const double per_sm_op = per_warp_op / (blocks_per_sm * threads / 32); // SM cycles per warp-level operation
    printf("%-28s warps/SM %2d keys %5d: %8.1f cycles per op per warp, %6.2f SM-cycles per op\n", name,
           blocks_per_sm * threads / 32, mask + 1, per_warp_op, per_sm_op);
    cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
    for (int bps : {1, 6, 12}) {
        for (int mask : {0, 15, 4095}) run<0>("match.any", bps, mask);
        run<1>("12 ballots", bps, 4095);
        run<2>("bitonic 15 steps + heads", bps, 4095);
        run<3>("one shuffle", bps, 4095);
    }
    return 0;
}
