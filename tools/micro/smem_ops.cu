// Micro-benchmark: shared-memory pipe cost of the operations the onesweep pass is built from, at the pass kernel's own
// occupancy (2 CTAs x 384 threads per SM, per-warp 256-entry counters, random 8-bit digits).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/smem_ops tools/micro/smem_ops.cu
// Prints SM cycles per warp instruction (both CTAs of an SM running) for every operation.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int THREADS = 384, WARPS = THREADS / 32, ITEMS = 24, REPS = 64;
constexpr int TILE = THREADS * ITEMS;

struct smem_t
{
    alignas(16) uint32_t kv[2 * TILE];
    uint32_t hist[WARPS][256];
    uint32_t dummy[WARPS][32];
};

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

template <int OP>
__global__ void __launch_bounds__(THREADS, 2) bench(uint32_t* out, long long* cycles)
{
    extern __shared__ __align__(128) unsigned char raw[];
    smem_t& sm = *reinterpret_cast<smem_t*>(raw);
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < WARPS * 256; i += THREADS) (&sm.hist[0][0])[i] = 0;
    for (int i = tid; i < 2 * TILE; i += THREADS) sm.kv[i] = i;
    uint32_t d[ITEMS], pos[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; j++)
    {
        const uint32_t h = mix(blockIdx.x * 7919u + tid * 31u + j * 104729u + 12345u);
        d[j] = h & 255u;             // random digit
        pos[j] = (h >> 8) % TILE;    // random regroup position
    }
    __syncthreads();
    uint32_t* my = sm.hist[warp];
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < REPS; r++)
    {
#pragma unroll
        for (int j = 0; j < ITEMS; j++)
        {
            if (OP == 0) atomicAdd(&my[d[j]], 1u);                                     // non-returning add (count step)
            if (OP == 1) acc += atomicAdd(&my[d[j]], 1u);                              // returning add, all lanes
            if (OP == 2) acc += my[d[j]];                                             // random load
            if (OP == 3) my[d[j]] = acc + j;                                          // random store
            if (OP == 4) { const uint32_t v = my[d[j]]; acc += v; __syncwarp(); my[d[j]] = v + 1; __syncwarp(); }   // load + store
            if (OP == 5) reinterpret_cast<uint2*>(sm.kv)[pos[j]] = make_uint2(acc, j);  // 64-bit scatter
            if (OP == 6) { sm.kv[pos[j]] = acc; sm.kv[TILE + pos[j]] = j; }              // two 32-bit scatters
            if (OP == 7) { const uint2 e = reinterpret_cast<const uint2*>(sm.kv)[j * THREADS + tid]; acc += e.x + e.y; }   // coalesced 64-bit load
            if (OP == 8) acc += sm.kv[warp * (ITEMS * 32) + lane + j * 32];            // coalesced 32-bit load
            if (OP == 9)                                                               // returning add by ~28 leaders, others to a dummy word
            {
                const bool leader = (lane & 7u) != (uint32_t) (j & 7);
                acc += atomicAdd(leader ? &my[d[j]] : &sm.dummy[warp][lane], 3u);
            }
            if (OP == 10) acc += atomicAdd(&my[(lane * 8u + j) & 255u], 1u);           // returning add, conflict-free banks
            if (OP == 11) atomicAdd(&my[(lane * 8u + j) & 255u], 1u);                  // non-returning add, conflict-free banks
            if (OP == 12) acc += __shfl_sync(0xffffffffu, acc, d[j] & 31u);           // shuffle
            if (OP == 13) acc += __popc(__ballot_sync(0xffffffffu, (d[j] >> (r & 7)) & 1u));   // ballot + popc
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * THREADS + tid] = acc + sm.kv[tid] + my[lane];
}

template <int OP>
void run(const char* name, uint32_t* out, long long* cyc, int ctas)
{
    cudaFuncSetAttribute(bench<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(smem_t));
    bench<OP><<<ctas, THREADS, sizeof(smem_t)>>>(out, cyc);
    bench<OP><<<ctas, THREADS, sizeof(smem_t)>>>(out, cyc);
    cudaDeviceSynchronize();
    static long long h[4096];
    cudaMemcpy(h, cyc, sizeof(long long) * ctas, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < ctas; i++) avg += (double) h[i];
    avg /= ctas;
    // both CTAs of an SM run for ~avg cycles and issue 2 * WARPS * ITEMS * REPS warp instructions of the operation
    const double per = avg / (2.0 * WARPS * ITEMS * REPS);
    printf("{\"op\": \"%s\", \"sm_cycles_per_warp_instruction\": %.2f, \"cta_cycles\": %.0f, \"err\": \"%s\"}\n", name, per, avg, cudaGetErrorString(cudaGetLastError()));
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms * 2;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * ctas * THREADS);
    cudaMalloc(&cyc, sizeof(long long) * ctas);
    run<0>("red.add random digit (count step)", out, cyc, ctas);
    run<1>("atom.add returning, random digit, all lanes", out, cyc, ctas);
    run<2>("ld random digit", out, cyc, ctas);
    run<3>("st random digit", out, cyc, ctas);
    run<4>("ld + st random digit (counter read-modify-write)", out, cyc, ctas);
    run<5>("st.64 scatter to random positions (regroup)", out, cyc, ctas);
    run<6>("2 x st.32 scatter to random positions", out, cyc, ctas);
    run<7>("ld.64 coalesced (write-out read-back)", out, cyc, ctas);
    run<8>("ld.32 coalesced (staged key load)", out, cyc, ctas);
    run<9>("atom.add returning, 28 leaders + 4 dummy lanes", out, cyc, ctas);
    run<10>("atom.add returning, conflict-free banks", out, cyc, ctas);
    run<11>("red.add conflict-free banks", out, cyc, ctas);
    run<12>("shfl.idx", out, cyc, ctas);
    run<13>("ballot + popc", out, cyc, ctas);
    return 0;
}
