// A C++ host — the reference's language — driving the multi-GPU radix sort through the C ABI alone (include/vrenb200.h,
// vrenb200_sharded_sort_*): no Python, no torch, no NCCL.  This is the program INTEGRATION.md sketches.  All ranks live in
// this process; with at least `world` GPUs that can address each other every rank gets its own device (peer access enabled),
// otherwise all ranks share device 0 (their "peers" are then ordinary allocations of that device — the same code path of the
// library, the way tests/run_sharded_emulated.py exercises it from Python).
//   sharded_sort_host <world> <pairs per rank> [<rounds>]
// The concatenation of the ranks' outputs must be the stable sort by key of the concatenation of their inputs
// (std::stable_sort, the reference test's check: vren_test radix_sort.cpp:88, key/value extension).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "vrenb200.h"

#define CUDA_OK(expr)                                                                             \
    do                                                                                            \
    {                                                                                             \
        cudaError_t e_ = (expr);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
        {                                                                                         \
            std::printf("FAIL %s:%d: %s -> %s\n", __FILE__, __LINE__, #expr, cudaGetErrorString(e_)); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)
#define VREN_OK(expr)                                                                             \
    do                                                                                            \
    {                                                                                             \
        int s_ = (expr);                                                                          \
        if (s_ != VRENB200_OK)                                                                    \
        {                                                                                         \
            std::printf("FAIL %s:%d: %s -> status %d\n", __FILE__, __LINE__, #expr, s_);           \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

int main(int argc, char** argv)
{
    // the ranks' kernels wait for each other's flags: on one device every stream needs its own hardware queue
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    const uint32_t world = argc > 1 ? (uint32_t) std::atoi(argv[1]) : 2;
    const uint32_t n = argc > 2 ? (uint32_t) std::strtoul(argv[2], nullptr, 10) : 200000;
    const uint32_t rounds = argc > 3 ? (uint32_t) std::atoi(argv[3]) : 2;
    if (world < 1 || world > 8) { std::printf("world must be 1..8\n"); return 2; }

    int ndev = 0;
    CUDA_OK(cudaGetDeviceCount(&ndev));
    bool one_per_gpu = ndev >= (int) world && world > 1;
    if (one_per_gpu)
        for (uint32_t a = 0; a < world && one_per_gpu; a++)
            for (uint32_t b = 0; b < world; b++)
            {
                int can = 1;
                if (a != b) CUDA_OK(cudaDeviceCanAccessPeer(&can, (int) a, (int) b));
                if (!can) one_per_gpu = false;
            }
    auto device_of = [&](uint32_t r) { return one_per_gpu ? (int) r : 0; };
    if (one_per_gpu)
        for (uint32_t a = 0; a < world; a++)
        {
            CUDA_OK(cudaSetDevice((int) a));
            for (uint32_t b = 0; b < world; b++)
                if (a != b)
                {
                    cudaError_t e = cudaDeviceEnablePeerAccess((int) b, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CUDA_OK(e);
                    cudaGetLastError();
                }
        }
    std::printf("%u ranks, %u pairs each, %u rounds, %s\n", world, n, rounds, one_per_gpu ? "one GPU per rank (peer access)" : "all ranks on device 0");

    // sizes: the receive capacity of include/vrenb200.h's rule of thumb (balanced keys)
    const uint32_t capacity = (uint32_t) ((double) n * 1.25) + 257u * 12288u;
    const size_t sym = vrenb200_sharded_sort_symmetric_bytes(capacity);
    const size_t local_bytes = vrenb200_sharded_sort_local_bytes(n, capacity, nullptr);
    std::vector<void*> region(world), local(world);
    std::vector<cudaStream_t> stream(world);
    std::vector<uint32_t*> dkeys(world), dvals(world);
    std::vector<vrenb200_sharded_sort*> ctx(world, nullptr);
    for (uint32_t r = 0; r < world; r++)
    {
        CUDA_OK(cudaSetDevice(device_of(r)));
        CUDA_OK(cudaMalloc(&region[r], sym));
        CUDA_OK(cudaMalloc(&local[r], local_bytes));
        CUDA_OK(cudaMalloc(&dkeys[r], (size_t) n * 4 + 16));
        CUDA_OK(cudaMalloc(&dvals[r], (size_t) n * 4 + 16));
        CUDA_OK(cudaStreamCreateWithFlags(&stream[r], cudaStreamNonBlocking));
    }
    for (uint32_t r = 0; r < world; r++)
    {
        CUDA_OK(cudaSetDevice(device_of(r)));
        VREN_OK(vrenb200_sharded_sort_create(&ctx[r], r, world, n, capacity, rounds, region.data(), local[r], local_bytes, nullptr));
    }
    for (uint32_t r = 0; r < world; r++)      // "synchronise the ranks once before the first sort"
    {
        CUDA_OK(cudaSetDevice(device_of(r)));
        CUDA_OK(cudaDeviceSynchronize());
    }

    // input: uniform keys, value = global index (unique, rank-major) — stability is visible in the values
    std::mt19937 rng(2024);
    std::vector<std::vector<uint32_t>> keys(world, std::vector<uint32_t>(n)), vals(world, std::vector<uint32_t>(n));
    std::vector<uint32_t> all_k, all_v;
    for (uint32_t r = 0; r < world; r++)
    {
        for (uint32_t i = 0; i < n; i++)
        {
            keys[r][i] = (r & 1u) ? (rng() & 0xFFFF00FFu) : rng();      // odd ranks: repeated keys
            vals[r][i] = r * n + i;
        }
        all_k.insert(all_k.end(), keys[r].begin(), keys[r].end());
        all_v.insert(all_v.end(), vals[r].begin(), vals[r].end());
        CUDA_OK(cudaSetDevice(device_of(r)));
        CUDA_OK(cudaMemcpy(dkeys[r], keys[r].data(), (size_t) n * 4, cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(dvals[r], vals[r].data(), (size_t) n * 4, cudaMemcpyHostToDevice));
    }
    std::vector<uint32_t> order(all_k.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t) i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return all_k[a] < all_k[b]; });

    for (int rep = 0; rep < 2; rep++)        // twice: the receive buffers, flags and epochs are reused
    {
        // every rank enqueues its part; nothing blocks the host until the streams are synchronised
        for (uint32_t r = 0; r < world; r++)
        {
            CUDA_OK(cudaSetDevice(device_of(r)));
            VREN_OK(vrenb200_sharded_sort_pairs(ctx[r], (vrenb200_stream_t) stream[r], dkeys[r], dvals[r], n, 32));
        }
        for (uint32_t r = 0; r < world; r++)
        {
            CUDA_OK(cudaSetDevice(device_of(r)));
            CUDA_OK(cudaStreamSynchronize(stream[r]));
        }
        std::vector<uint32_t> got_k, got_v;
        for (uint32_t r = 0; r < world; r++)
        {
            CUDA_OK(cudaSetDevice(device_of(r)));
            uint32_t status[8];
            CUDA_OK(cudaMemcpy(status, vrenb200_sharded_sort_status(ctx[r]), sizeof(status), cudaMemcpyDeviceToHost));
            if (status[0] != 0) { std::printf("FAIL rank %u: the plan does not fit (status %u)\n", r, status[0]); return 1; }
            const uint32_t m = status[1];
            std::vector<uint32_t> k(m), v(m);
            CUDA_OK(cudaMemcpy(k.data(), vrenb200_sharded_sort_out_keys(ctx[r]), (size_t) m * 4, cudaMemcpyDeviceToHost));
            CUDA_OK(cudaMemcpy(v.data(), vrenb200_sharded_sort_out_values(ctx[r]), (size_t) m * 4, cudaMemcpyDeviceToHost));
            std::printf("  rank %u: %u pairs, digits [%u, %u) of byte %u\n", r, m, status[2], status[3], status[4]);
            got_k.insert(got_k.end(), k.begin(), k.end());
            got_v.insert(got_v.end(), v.begin(), v.end());
        }
        if (got_k.size() != all_k.size()) { std::printf("FAIL: %zu pairs out, %zu in\n", got_k.size(), all_k.size()); return 1; }
        for (size_t i = 0; i < order.size(); i++)
            if (got_k[i] != all_k[order[i]] || got_v[i] != all_v[order[i]])
            {
                std::printf("FAIL rep %d: pair %zu is (%08x, %u), expected (%08x, %u)\n", rep, i, got_k[i], got_v[i], all_k[order[i]], all_v[order[i]]);
                return 1;
            }
    }
    for (uint32_t r = 0; r < world; r++)
    {
        CUDA_OK(cudaSetDevice(device_of(r)));
        vrenb200_sharded_sort_destroy(ctx[r]);
        cudaFree(region[r]); cudaFree(local[r]); cudaFree(dkeys[r]); cudaFree(dvals[r]);
        cudaStreamDestroy(stream[r]);
    }
    std::printf("ALL PASS\n");
    return 0;
}
