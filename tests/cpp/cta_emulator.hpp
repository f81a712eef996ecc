// Host execution of a CUDA kernel BODY, thread for thread: one OS thread per CUDA thread of ONE CTA, __syncthreads() and the
// warp-synchronous intrinsics built on pthread barriers.  A body written against threadIdx.x, __syncthreads, __syncthreads_count,
// __syncwarp, __shfl_sync, __shfl_up_sync, __match_any_sync, __ballot_sync, __reduce_add_sync, atomicOr / atomicAdd, __shared__
// variables, blockIdx.x (CTAs of a grid run one after the other), __popc, __clz, __ffs (full masks, convergent warps) compiles
// unchanged with g++
// when this header is included BEFORE it.  Built with -fsanitize=thread, a missing barrier between two accesses of the same
// shared word shows up as a data race (pthread barriers are synchronisation ThreadSanitizer understands) — the host-side
// stand-in for compute-sanitizer's racecheck when no GPU is at hand.
// Test infrastructure only (tests/test_small_sort_emulation.py); nothing of the product includes it.
#pragma once

#include <pthread.h>
#include <sched.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <vector>

// (a translation unit that needs the CUDA runtime's host types includes <cuda_runtime.h> BEFORE this header: its host_defines.h
// gives these qualifiers attribute meanings that g++ does not know)
#undef __device__
#undef __global__
#undef __forceinline__
#undef __shared__
#define __device__
#define __global__
#define __forceinline__ inline
#define __shared__ static      // one CTA at a time: a function-level static IS the CTA's shared variable

struct emu_uint3 { unsigned x, y, z; };
static thread_local emu_uint3 threadIdx = {0, 0, 0};
static thread_local emu_uint3 blockIdx = {0, 0, 0};      // set by run_cta's `block` argument: CTAs of a grid run one after the other

namespace cta_emu {

struct warp_state
{
    pthread_barrier_t bar;
    uint64_t slot[32];
};

struct cta_state
{
    pthread_barrier_t bar;
    std::vector<warp_state> warps;
    int vote = 0;               // __syncthreads_count
};

static thread_local cta_state* cta = nullptr;

inline warp_state& my_warp() { return cta->warps[threadIdx.x >> 5]; }
inline unsigned my_lane() { return threadIdx.x & 31u; }

// threads start behind a gate, so that a host that cannot create all of them (thread limit of a container) is reported
// instead of leaving the ones that did start waiting at their first barrier
struct start_gate
{
    pthread_mutex_t mutex = PTHREAD_MUTEX_INITIALIZER;
    pthread_cond_t cond = PTHREAD_COND_INITIALIZER;
    int state = 0;      // 0 wait, 1 run the body, 2 give up
};

struct thread_arg
{
    cta_state* cta;
    unsigned tid;
    unsigned block;
    const std::function<void()>* body;
    start_gate* gate;
};

inline void* thread_main(void* p)
{
    thread_arg* a = static_cast<thread_arg*>(p);
    pthread_mutex_lock(&a->gate->mutex);
    while (a->gate->state == 0) pthread_cond_wait(&a->gate->cond, &a->gate->mutex);
    const int state = a->gate->state;
    pthread_mutex_unlock(&a->gate->mutex);
    if (state != 1) return nullptr;
    threadIdx.x = a->tid;
    blockIdx.x = a->block;
    cta = a->cta;
    (*a->body)();
    return nullptr;
}

// runs `body` once per thread of a CTA of `threads` threads (a multiple of 32); false: the host could not create the threads
inline bool run_cta(unsigned threads, const std::function<void()>& body, unsigned block = 0)
{
    cta_state st;
    st.warps.resize(threads / 32);
    pthread_barrier_init(&st.bar, nullptr, threads);
    for (auto& w : st.warps) pthread_barrier_init(&w.bar, nullptr, 32);
    std::vector<pthread_t> ids(threads);
    std::vector<thread_arg> args(threads);
    start_gate gate;
    pthread_attr_t attr;
    pthread_attr_init(&attr);
    pthread_attr_setstacksize(&attr, 512 * 1024);
    unsigned created = 0;
    for (; created < threads; created++)
    {
        args[created] = thread_arg{&st, created, block, &body, &gate};
        if (pthread_create(&ids[created], &attr, thread_main, &args[created]) != 0) break;
    }
    pthread_mutex_lock(&gate.mutex);
    gate.state = created == threads ? 1 : 2;
    pthread_cond_broadcast(&gate.cond);
    pthread_mutex_unlock(&gate.mutex);
    for (unsigned t = 0; t < created; t++) pthread_join(ids[t], nullptr);
    pthread_attr_destroy(&attr);
    pthread_barrier_destroy(&st.bar);
    for (auto& w : st.warps) pthread_barrier_destroy(&w.bar);
    if (created != threads) std::fprintf(stderr, "cta_emulator: this host allows only %u of the %u threads of the CTA\n", created, threads);
    return created == threads;
}

} // namespace cta_emu

inline void __syncthreads() { pthread_barrier_wait(&cta_emu::cta->bar); }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { pthread_barrier_wait(&cta_emu::my_warp().bar); }

// number of threads of the CTA whose predicate is non-zero; a barrier like __syncthreads()
inline int __syncthreads_count(int pred)
{
    cta_emu::cta_state* c = cta_emu::cta;
    if (pred) __atomic_fetch_add(&c->vote, 1, __ATOMIC_RELAXED);
    pthread_barrier_wait(&c->bar);
    const int r = __atomic_load_n(&c->vote, __ATOMIC_RELAXED);
    pthread_barrier_wait(&c->bar);
    if (threadIdx.x == 0) __atomic_store_n(&c->vote, 0, __ATOMIC_RELAXED);
    pthread_barrier_wait(&c->bar);
    return r;
}

inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

template <typename T>
inline T __shfl_sync(unsigned, T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    cta_emu::warp_state& w = cta_emu::my_warp();
    w.slot[cta_emu::my_lane()] = (uint64_t) v;
    pthread_barrier_wait(&w.bar);
    const T r = (T) w.slot[(unsigned) src_lane & 31u];
    pthread_barrier_wait(&w.bar);
    return r;
}

template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    cta_emu::warp_state& w = cta_emu::my_warp();
    const unsigned lane = cta_emu::my_lane();
    w.slot[lane] = (uint64_t) v;
    pthread_barrier_wait(&w.bar);
    const T r = lane >= delta ? (T) w.slot[lane - delta] : v;
    pthread_barrier_wait(&w.bar);
    return r;
}

inline unsigned __match_any_sync(unsigned, uint32_t v)
{
    cta_emu::warp_state& w = cta_emu::my_warp();
    w.slot[cta_emu::my_lane()] = v;
    pthread_barrier_wait(&w.bar);
    unsigned m = 0;
    for (unsigned i = 0; i < 32; i++)
        if (w.slot[i] == (uint64_t) v) m |= 1u << i;
    pthread_barrier_wait(&w.bar);
    return m;
}

inline unsigned __ballot_sync(unsigned, int pred)
{
    cta_emu::warp_state& w = cta_emu::my_warp();
    w.slot[cta_emu::my_lane()] = pred ? 1u : 0u;
    pthread_barrier_wait(&w.bar);
    unsigned m = 0;
    for (unsigned i = 0; i < 32; i++)
        if (w.slot[i]) m |= 1u << i;
    pthread_barrier_wait(&w.bar);
    return m;
}

inline uint32_t __reduce_add_sync(unsigned, uint32_t v)
{
    cta_emu::warp_state& w = cta_emu::my_warp();
    w.slot[cta_emu::my_lane()] = v;
    pthread_barrier_wait(&w.bar);
    uint32_t sum = 0;
    for (unsigned i = 0; i < 32; i++) sum += (uint32_t) w.slot[i];
    pthread_barrier_wait(&w.bar);
    return sum;
}

inline int __ffs(int x) { return __builtin_ffs(x); }
inline void __nanosleep(unsigned) { sched_yield(); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned) x); }
