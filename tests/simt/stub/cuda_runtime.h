// TEST INFRASTRUCTURE ONLY -- host-side SIMT emulator standing in for <cuda_runtime.h> when the kernel sources
// of differentialriccatiequations.jl_b200/csrc are compiled with g++ -DDRE_SIMT_EMU for the "not gpu" test tier
// (tests/simt/).  The product library (libdre_b200.so) is never built from this header and never links it.
//
// Execution model: one CTA at a time; every CUDA thread is a ucontext fiber that runs until it blocks at a
// CTA barrier (__syncthreads) or a warp collective (__shfl_*_sync, __syncwarp, the m8n8k4 DMMA) or returns.
// Fibers run in thread order, i.e. with the largest possible skew between threads, so a missing barrier shows
// up as a wrong result; shared memory is filled with NaN patterns before each CTA; cp.async copies are
// deferred until the matching wait_group.  Kernel launches are synchronous; streams are ignored.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <mutex>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct double2 { double x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

// ---- runtime API subset used by csrc/*.cu: synchronous host implementations ----
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
typedef void* cudaMemPool_t;
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
struct cudaDeviceProp { int major, minor, multiProcessorCount; char name[64]; };
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "simt emulator: no error text"; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; strcpy(p->name, "SIMT emulator (host)");
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned, int) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (void*)1; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
// events carry a host time stamp (everything is synchronous, so "recorded" == "completed")
static inline double simt_now_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = calloc(1, sizeof(double)); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { *(double*)e = simt_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(*(double*)b - *(double*)a);
    return cudaSuccess;
}
// "device" memory is host memory filled with NaN patterns (the kernels must define what they read)
static inline cudaError_t cudaMalloc(void** p, size_t n) {
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    if (n <= ((size_t)64 << 20)) memset(*p, 0xFF, n);   // (arena chunks of several GB stay untouched / uncommitted)
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width,
                                            size_t height, cudaMemcpyKind, cudaStream_t) {
    for (size_t i = 0; i < height; ++i) memmove((char*)d + i * dpitch, (const char*)s + i * spitch, width);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }

using std::isinf;
using std::isnan;
using std::max;
using std::min;
static inline int64_t min(int64_t a, int b) { return a < b ? a : (int64_t)b; }
static inline int64_t min(int a, int64_t b) { return a < b ? (int64_t)a : b; }
static inline int64_t max(int64_t a, int b) { return a > b ? a : (int64_t)b; }
static inline int64_t max(int a, int64_t b) { return a > b ? (int64_t)a : b; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }

// Context switch: a dozen instructions on x86-64 (callee-saved registers + stack pointer); the ucontext
// fallback costs a sigprocmask system call per switch.
#if defined(__x86_64__)
#define SIMT_FAST_SWITCH 1
extern "C" void simt_switch(void** save_sp, void* load_sp);
#ifdef SIMT_EMU_IMPL
asm(".text\n.globl simt_switch\n.hidden simt_switch\n.type simt_switch,@function\nsimt_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size simt_switch,.-simt_switch\n");
#endif
#endif

namespace simt {

enum State { RUN = 0, BLOCK_CTA, BLOCK_WARP, DONE };

struct PendingCopy { void* dst; const void* src; int bytes, src_bytes; };

struct Fiber {
    ucontext_t ctx;
    void* sp = nullptr;
    char* stack = nullptr;
    State state = DONE;
    uint3 tid{0, 0, 0};
    int lin = 0;
    std::vector<PendingCopy> cp;        // deferred cp.async copies
    std::vector<size_t> cp_groups;      // end offsets (into cp) of the committed groups
};

struct WarpState {
    alignas(16) unsigned char in[32][64];
    alignas(16) unsigned char out[32][64];
    bool present[32];
    int arrived = 0, alive = 0;
};

struct Machine {
    static constexpr size_t STACK = 256 * 1024;
    static constexpr size_t DYN_SMEM = 232448;
    ucontext_t sched;
    void* sched_sp = nullptr;
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    Fiber* cur = nullptr;
    int nthreads = 0, cta_alive = 0, cta_arrived = 0;
    dim3 grid, block;
    uint3 bid{0, 0, 0};
    alignas(16) unsigned char dyn_smem[DYN_SMEM];
    std::function<void()> body;
    long launches = 0, ctas = 0, switches = 0;
};

inline Machine& M() {
    static Machine* m = new Machine();
    return *m;
}

inline void die(const char* msg) {
    fprintf(stderr, "simt emulator: %s\n", msg);
    abort();
}

inline void yield_to_scheduler() {
    Machine& m = M();
    m.switches++;
#ifdef SIMT_FAST_SWITCH
    simt_switch(&m.cur->sp, m.sched_sp);
#else
    swapcontext(&m.cur->ctx, &m.sched);
#endif
}

inline void release_cta_if_complete() {
    Machine& m = M();
    if (m.cta_arrived > 0 && m.cta_arrived == m.cta_alive) {
        m.cta_arrived = 0;
        for (int i = 0; i < m.nthreads; ++i)
            if (m.fibers[i].state == BLOCK_CTA) m.fibers[i].state = RUN;
    }
}

inline void cta_barrier() {
    Machine& m = M();
    m.cta_arrived++;
    m.cur->state = BLOCK_CTA;
    release_cta_if_complete();
    if (m.cur->state == BLOCK_CTA) yield_to_scheduler();
}

// generic warp collective: every live lane deposits `in`; the last one to arrive runs all(in[], out[], present[])
template <class In, class Out, class F>
inline Out collective(const In& in, F all) {
    static_assert(sizeof(In) <= 64 && sizeof(Out) <= 64, "collective payload too large");
    Machine& m = M();
    const int lane = m.cur->lin & 31;
    WarpState& w = m.warps[m.cur->lin >> 5];
    memcpy(w.in[lane], &in, sizeof(In));
    w.present[lane] = true;
    w.arrived++;
    if (w.arrived == w.alive) {
        In ins[32];
        Out outs[32];
        for (int l = 0; l < 32; ++l)
            if (w.present[l]) memcpy(&ins[l], w.in[l], sizeof(In));
            else memset(&ins[l], 0, sizeof(In));
        all(ins, outs, w.present);
        const int base = (m.cur->lin >> 5) << 5;
        for (int l = 0; l < 32; ++l) {
            memcpy(w.out[l], &outs[l], sizeof(Out));
            if (w.present[l] && base + l < m.nthreads && m.fibers[base + l].state == BLOCK_WARP)
                m.fibers[base + l].state = RUN;
            w.present[l] = false;
        }
        w.arrived = 0;
    } else {
        m.cur->state = BLOCK_WARP;
        yield_to_scheduler();
    }
    Out o;
    memcpy(&o, w.out[lane], sizeof(Out));
    return o;
}

inline void fiber_entry() {
    Machine& m = M();
    m.body();
    // thread exit: it no longer takes part in barriers / collectives
    Fiber* f = m.cur;
    if (!f->cp.empty()) {
        for (auto& c : f->cp) {
            memset(c.dst, 0, c.bytes);
            memcpy(c.dst, c.src, c.src_bytes);
        }
        f->cp.clear();
        f->cp_groups.clear();
    }
    f->state = DONE;
    m.cta_alive--;
    WarpState& w = m.warps[f->lin >> 5];
    w.alive--;
    if (w.arrived > 0 && w.arrived == w.alive) die("a lane exited while the rest of its warp waits in a collective");
    release_cta_if_complete();
#ifdef SIMT_FAST_SWITCH
    simt_switch(&f->sp, m.sched_sp);
#else
    swapcontext(&f->ctx, &m.sched);
#endif
    die("resumed a finished fiber");
}

inline void run_cta() {
    Machine& m = M();
    const int nt = m.nthreads;
    if ((int)m.fibers.size() < nt) m.fibers.resize(nt);
    m.warps.assign((nt + 31) / 32, WarpState());
    for (auto& w : m.warps) {
        memset(w.present, 0, sizeof(w.present));
    }
    memset(m.dyn_smem, 0xFF, Machine::DYN_SMEM);
    m.cta_alive = nt;
    m.cta_arrived = 0;
    for (int i = 0; i < nt; ++i) {
        Fiber& f = m.fibers[i];
        if (!f.stack) {
            f.stack = (char*)mmap(nullptr, Machine::STACK, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            if (f.stack == MAP_FAILED) die("mmap of a fiber stack failed");
        }
        f.lin = i;
        f.tid.x = i % m.block.x;
        f.tid.y = (i / m.block.x) % m.block.y;
        f.tid.z = i / (m.block.x * m.block.y);
        f.state = RUN;
        f.cp.clear();
        f.cp_groups.clear();
        m.warps[i >> 5].alive++;
#ifdef SIMT_FAST_SWITCH
        void** top = (void**)(f.stack + Machine::STACK);   // 16-byte aligned
        top[-1] = nullptr;                                  // fake return address of fiber_entry
        top[-2] = (void*)fiber_entry;                       // popped by the `ret` of the first switch
        for (int k = 3; k <= 8; ++k) top[-k] = nullptr;     // rbp rbx r12 r13 r14 r15
        f.sp = (void*)(top - 8);
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = Machine::STACK;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
#endif
    }
    // Scheduling order of the runnable fibers: thread order by default (maximal skew in one direction);
    // SIMT_SCHED=reverse / random exposes dependences on the opposite / an arbitrary interleaving.
    static const int sched_mode = [] {
        const char* ev = getenv("SIMT_SCHED");
        return !ev ? 0 : (strcmp(ev, "reverse") == 0 ? 1 : (strcmp(ev, "random") == 0 ? 2 : 0));
    }();
    static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
    std::vector<int> order(nt);
    for (int i = 0; i < nt; ++i) order[i] = sched_mode == 1 ? nt - 1 - i : i;
    while (m.cta_alive > 0) {
        bool progress = false;
        if (sched_mode == 2)
            for (int i = nt - 1; i > 0; --i) {
                rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
                std::swap(order[i], order[(int)((rng_state >> 33) % (uint64_t)(i + 1))]);
            }
        for (int oi = 0; oi < nt; ++oi) {
            const int i = order[oi];
            Fiber& f = m.fibers[i];
            if (f.state != RUN) continue;
            m.cur = &f;
            progress = true;
#ifdef SIMT_FAST_SWITCH
            simt_switch(&m.sched_sp, f.sp);
#else
            swapcontext(&m.sched, &f.ctx);
#endif
        }
        if (!progress) die("deadlock: every live thread of the CTA is blocked (divergent barrier / collective?)");
    }
    m.cur = nullptr;
    m.ctas++;
}

inline std::mutex& launch_mutex() {
    static std::mutex mx;
    return mx;
}

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& body) {
    std::lock_guard<std::mutex> lock(launch_mutex());   // one kernel at a time, whichever host thread launches it
    Machine& m = M();
    if (m.cur) die("nested kernel launch");
    if (smem > Machine::DYN_SMEM) die("dynamic shared memory request exceeds 227 KB");
    m.grid = grid;
    m.block = block;
    m.nthreads = (int)(block.x * block.y * block.z);
    if (m.nthreads <= 0 || m.nthreads > 1024) die("bad block size");
    m.body = std::forward<F>(body);
    m.launches++;
    // CTAs run one after the other: in launch order by default, reversed / shuffled with SIMT_SCHED=reverse / random
    // (a kernel whose CTAs depend on each other's results within one launch is wrong on the GPU)
    const char* ev = getenv("SIMT_SCHED");
    const int mode = !ev ? 0 : (strcmp(ev, "reverse") == 0 ? 1 : (strcmp(ev, "random") == 0 ? 2 : 0));
    const uint64_t total = (uint64_t)grid.x * grid.y * grid.z;
    std::vector<uint64_t> order;
    if (mode == 2) {
        order.resize(total);
        for (uint64_t i = 0; i < total; ++i) order[i] = i;
        uint64_t st = 0x2545F4914F6CDD1Dull ^ total;
        for (uint64_t i = total; i > 1; --i) {
            st = st * 6364136223846793005ull + 1442695040888963407ull;
            std::swap(order[i - 1], order[(st >> 33) % i]);
        }
    }
    for (uint64_t k = 0; k < total; ++k) {
        const uint64_t lin = mode == 2 ? order[k] : (mode == 1 ? total - 1 - k : k);
        m.bid = uint3{(unsigned)(lin % grid.x), (unsigned)((lin / grid.x) % grid.y), (unsigned)(lin / ((uint64_t)grid.x * grid.y))};
        run_cta();
    }
    m.body = nullptr;
}

// ---- DMMA m8n8k4: lane l holds a = A[l/4][l%4], b = B[l%4][l/4], c0,c1 = C[l/4][2*(l%4) + {0,1}] ----
struct DmmaIn { double a, b, c0, c1; };
struct DmmaOut { double c0, c1; };
inline void dmma884(double& c0, double& c1, double a, double b) {
    DmmaOut o = collective<DmmaIn, DmmaOut>(DmmaIn{a, b, c0, c1}, [](const DmmaIn* in, DmmaOut* out, const bool* present) {
        for (int l = 0; l < 32; ++l)
            if (!present[l]) die("mma.sync executed by a partial warp");
        for (int l = 0; l < 32; ++l) {
            const int row = l >> 2, col = 2 * (l & 3);
            double s0 = in[l].c0, s1 = in[l].c1;
            for (int k = 0; k < 4; ++k) {
                const double av = in[row * 4 + k].a;
                s0 = std::fma(av, in[col * 4 + k].b, s0);
                s1 = std::fma(av, in[(col + 1) * 4 + k].b, s1);
            }
            out[l].c0 = s0;
            out[l].c1 = s1;
        }
    });
    c0 = o.c0;
    c1 = o.c1;
}

template <class T, class Pick>
inline T shuffle(T v, Pick pick) {
    struct In { T v; int src; };
    const int lane = M().cur->lin & 31;
    In in{v, pick(lane)};
    return collective<In, T>(in, [](const In* ins, T* outs, const bool* present) {
        for (int l = 0; l < 32; ++l) {
            const int s = ins[l].src;
            outs[l] = (s >= 0 && s < 32 && present[s]) ? ins[s].v : ins[l].v;
        }
    });
}

// ---- cp.async (deferred until the matching wait_group) ----
inline void cp_async(void* dst, const void* src, int bytes, int src_bytes) {
    M().cur->cp.push_back(PendingCopy{dst, src, bytes, src_bytes});
}
inline void cp_async_commit() { M().cur->cp_groups.push_back(M().cur->cp.size()); }
inline void cp_async_wait(int keep) {
    Fiber* f = M().cur;
    const int ngroups = (int)f->cp_groups.size();
    const int done = ngroups - keep;
    if (done <= 0) return;
    const size_t upto = f->cp_groups[done - 1];
    for (size_t i = 0; i < upto; ++i) {
        const PendingCopy& c = f->cp[i];
        memset(c.dst, 0, c.bytes);
        if (c.src_bytes > 0) memcpy(c.dst, c.src, c.src_bytes);
    }
    f->cp.erase(f->cp.begin(), f->cp.begin() + upto);
    f->cp_groups.erase(f->cp_groups.begin(), f->cp_groups.begin() + done);
    for (auto& g : f->cp_groups) g -= upto;
}

}  // namespace simt

#define threadIdx (simt::M().cur->tid)
#define blockIdx (simt::M().bid)
#define blockDim (simt::M().block)
#define gridDim (simt::M().grid)

static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    simt::collective<int, int>(0, [](const int*, int* o, const bool*) { for (int l = 0; l < 32; ++l) o[l] = 0; });
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    return simt::shuffle<T>(v, [=](int lane) { return (lane & ~(width - 1)) + (src & (width - 1)); });
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int mask, int width = 32) {
    return simt::shuffle<T>(v, [=](int lane) {
        const int s = lane ^ mask;
        return (s & ~(width - 1)) == (lane & ~(width - 1)) ? s : lane;
    });
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    return simt::shuffle<T>(v, [=](int lane) {
        const int s = lane + (int)delta;
        return (s & ~(width - 1)) == (lane & ~(width - 1)) ? s : lane;
    });
}
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }

// kernel launch and dynamic shared memory as the kernel sources spell them (see csrc/common.cuh)
#define DRE_LAUNCH(kernel, grid, block, smem, stream, ...) \
    simt::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { DRE_UNPAREN kernel(__VA_ARGS__); })
#define DRE_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(simt::M().dyn_smem)
#define DRE_DYN_SMEM_ALIGNED(type, name) DRE_DYN_SMEM(type, name)
