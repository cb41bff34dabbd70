// TEST INFRASTRUCTURE ONLY -- never part of the product path, never shipped in liblexp_cuda.so.
//
// A host-side emulation of the slice of the CUDA programming model that localexpstereo_b200/csrc uses, so that the SAME
// kernel source (lexp_kernels.cuh) and the SAME host code (lexp_capi.cu: planner, tiling, work items, copies) can be
// compiled with g++ -DLEXP_EMU into tests/emu/liblexp_emu.so and exercised by the CPU test-suite:
//   * every CUDA thread of a block is a fiber (ucontext); fibers switch only at barriers, so a block is executed as a
//     deterministic interleaving that the scheduler can permute (LEXP_EMU_ORDER = 0 forward, 1 reverse, 2 shuffled per
//     pass): a missing barrier shows up as an order-dependent result, a wrong barrier count as a reported deadlock;
//   * named barriers (bar.sync / bar.arrive with a thread count) and __syncthreads() follow the PTX semantics;
//   * dynamic / static shared memory is one buffer per block (blocks run one after the other);
//   * "device memory" is host memory, streams are synchronous, device properties mimic a B200 (148 SMs, 227 KB opt-in
//     shared memory) so that the planner takes the decisions it takes on the real device.
// It answers "is the kernel's logic right" (indices, pipeline protocol, arithmetic within FP32 rounding) without a GPU.
// It says nothing about performance, memory-model subtleties below barrier granularity, or FMA contraction choices
// (results can differ from the GPU's in the last bits; parity tests use the same 1e-4 tolerance on both).
#pragma once
#ifndef LEXP_EMU
#error "tests/emu/cuda_runtime.h is only for -DLEXP_EMU builds of the test emulator"
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <unistd.h>

#include <algorithm>
#include <functional>
#include <mutex>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

// ---- qualifiers ---------------------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static thread_local  /* static shared arrays: one block runs at a time per OS thread */

// ---- vector types ---------------------------------------------------------------------------------------------------------
struct uchar4 { unsigned char x, y, z, w; };
struct uint2 { unsigned int x, y; };
struct int2 { int x, y; };
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

// ---- the fiber scheduler ------------------------------------------------------------------------------------------------
namespace emu {

struct Fiber {
    ucontext_t ctx;
    dim3 tid;
    bool done = false;
    int wait_bar = -1;       // barrier this fiber sleeps on, -1: runnable
    unsigned wait_gen = 0;
};
struct Barrier {
    int arrived = 0;
    unsigned gen = 0;
};
struct State {
    ucontext_t sched;
    Fiber* cur = nullptr;
    dim3 bid, bdim, gdim;
    Barrier bars[16];
    std::vector<unsigned char> dyn;      // dynamic shared memory of the running block
    std::vector<Fiber> fibers;
    std::vector<char> stacks;
    std::function<void()> entry;
    std::string error;
};
inline State& S() { static thread_local State s; return s; }
constexpr size_t kStack = 96 * 1024;

inline void yield_to_scheduler() { State& s = S(); swapcontext(&s.cur->ctx, &s.sched); }

// PTX bar.sync a, b / bar.arrive a, b: barrier a completes when b threads have arrived (by sync or arrive)
inline void barrier(int id, int expected, bool blocking) {
    State& s = S();
    Barrier& b = s.bars[id & 15];
    const unsigned my_gen = b.gen;
    if (++b.arrived == expected) {
        b.arrived = 0;
        b.gen++;
        return;
    }
    if (b.arrived > expected) { s.error = "barrier " + std::to_string(id) + ": more arrivals than its thread count " + std::to_string(expected); b.arrived = 0; b.gen++; return; }
    if (!blocking) return;
    s.cur->wait_bar = id & 15;
    s.cur->wait_gen = my_gen;
    yield_to_scheduler();
}

inline void trampoline() {
    State& s = S();
    s.entry();
    s.cur->done = true;
    yield_to_scheduler();
}

inline int order_mode() { const char* e = getenv("LEXP_EMU_ORDER"); return e ? atoi(e) : 0; }

// run one block: all fibers until completion
inline void run_block(unsigned nthreads) {
    State& s = S();
    if (s.fibers.size() < nthreads) { s.fibers.resize(nthreads); }
    if (s.stacks.size() < (size_t)nthreads * kStack) s.stacks.resize((size_t)nthreads * kStack);
    for (int i = 0; i < 16; i++) s.bars[i] = Barrier();
    for (unsigned i = 0; i < nthreads; i++) {
        Fiber& f = s.fibers[i];
        f.done = false; f.wait_bar = -1;
        f.tid = dim3(i % s.bdim.x, (i / s.bdim.x) % s.bdim.y, i / (s.bdim.x * s.bdim.y));
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = s.stacks.data() + (size_t)i * kStack;
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    std::vector<unsigned> order(nthreads);
    for (unsigned i = 0; i < nthreads; i++) order[i] = i;
    const int mode = order_mode();
    if (mode == 1) std::reverse(order.begin(), order.end());
    uint64_t lcg = 0x9E3779B97F4A7C15ull ^ ((uint64_t)s.bid.x * 2654435761u);
    unsigned remaining = nthreads;
    while (remaining) {
        if (mode == 2)
            for (unsigned i = nthreads - 1; i > 0; i--) { lcg = lcg * 6364136223846793005ull + 1442695040888963407ull; std::swap(order[i], order[(lcg >> 33) % (i + 1)]); }
        bool progress = false;
        for (unsigned k = 0; k < nthreads; k++) {
            Fiber& f = s.fibers[order[k]];
            if (f.done) continue;
            if (f.wait_bar >= 0) {
                if (s.bars[f.wait_bar].gen == f.wait_gen) continue;  // still asleep
                f.wait_bar = -1;
            }
            s.cur = &f;
            swapcontext(&s.sched, &f.ctx);
            progress = true;
            if (f.done) remaining--;
        }
        if (!progress) {
            std::string w;
            int shown = 0;
            for (unsigned i = 0; i < nthreads && shown < 6; i++)
                if (!s.fibers[i].done) { w += " t" + std::to_string(i) + "@bar" + std::to_string(s.fibers[i].wait_bar); shown++; }
            s.error = "deadlock in block " + std::to_string(s.bid.x) + ": " + std::to_string(remaining) + " threads asleep:" + w;
            return;
        }
    }
    for (int i = 0; i < 16; i++)
        if (s.bars[i].arrived != 0 && s.error.empty()) s.error = "block " + std::to_string(s.bid.x) + " ended with barrier " + std::to_string(i) + " half full";
}

inline std::string& last_launch_error() { static thread_local std::string e; return e; }

template <class... KArgs, class... Args>
inline void launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
#ifdef LEXP_EMU_NO_KERNELS  // host-logic-only builds (ThreadSanitizer on the C-ABI's locking: fibers and TSan do not mix)
    (void)kern; (void)grid; (void)block; (void)smem;
    usleep(300);  // a launch takes a while, so that concurrent callers really queue up behind it
    return;
#endif
    State& s = S();
    s.gdim = grid; s.bdim = block;
    s.error.clear();
#ifdef __SANITIZE_ADDRESS__
    s.dyn.assign(smem, 0xCD);       // exact size (malloc is 16-byte aligned): overruns of the dynamic shared memory are flagged
#else
    s.dyn.assign(smem + 64, 0xCD);  // poison: the kernel must initialise what it reads
#endif
    s.entry = [=]() { kern(args...); };
    const unsigned nthreads = block.x * block.y * block.z;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                s.bid = dim3(bx, by, bz);
                run_block(nthreads);
                if (!s.error.empty()) { last_launch_error() = s.error; return; }
            }
}

inline unsigned char* dyn_smem() { State& s = S(); return (unsigned char*)(((uintptr_t)s.dyn.data() + 15) & ~(uintptr_t)15); }

}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

static inline void __syncthreads() { emu::State& s = emu::S(); emu::barrier(0, (int)(s.bdim.x * s.bdim.y * s.bdim.z), true); }

// ---- device intrinsics ----------------------------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline int __double2int_rn(double v) { return (int)nearbyint(v); }
static inline int __float2int_rz(float v) { return (int)v; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) r |= (unsigned)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
}
static inline int atomicOr(int* p, int v) { int o = *p; *p |= v; return o; }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_ACQ_REL); }
static inline int atomicExch(int* p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_ACQ_REL); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __syncwarp(unsigned = 0xffffffffu) {}   /* fibers of a block switch only at barriers */
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline long long clock64() { return (long long)__builtin_ia32_rdtsc(); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t max(size_t a, size_t b) { return a > b ? a : b; }

// ---- runtime API (host memory stands in for device memory; streams are synchronous) -----------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorLaunchFailure = 719 };
typedef struct emu_stream* cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaHostRegisterPortable = 1, cudaHostRegisterMapped = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaLimit { cudaLimitPersistingL2CacheSize = 6 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
enum cudaAccessProperty { cudaAccessPropertyNormal = 0, cudaAccessPropertyStreaming = 1, cudaAccessPropertyPersisting = 2 };
enum cudaStreamAttrID { cudaStreamAttributeAccessPolicyWindow = 1 };
struct cudaAccessPolicyWindow { void* base_ptr; size_t num_bytes; float hitRatio; cudaAccessProperty hitProp, missProp; };
union cudaStreamAttrValue { cudaAccessPolicyWindow accessPolicyWindow; int syncPolicy; };
struct cudaPointerAttributes { cudaMemoryType type; int device; void* devicePointer; void* hostPointer; };
struct cudaDeviceProp {
    char name[256];
    int major, minor, multiProcessorCount, l2CacheSize, persistingL2CacheMaxSize, accessPolicyMaxWindowSize;
    size_t sharedMemPerBlockOptin, totalGlobalMem;
};

namespace emu {
inline std::string& reported_error();
inline std::map<void*, size_t>& registered() { static std::map<void*, size_t> s; return s; }
inline std::mutex& reg_mu() { static std::mutex m; return m; }
}  // namespace emu

static inline const char* cudaGetErrorString(cudaError_t e) {
    static thread_local std::string s;
    const std::string& why = emu::last_launch_error().empty() ? emu::reported_error() : emu::last_launch_error();
    s = e == cudaSuccess ? "no error" : ("emulated CUDA error " + std::to_string(e) + (why.empty() ? "" : ": " + why));
    return s.c_str();
}
// like CUDA: returns the pending launch error and resets it (the message stays readable through cudaGetErrorString)
inline std::string& emu::reported_error() { static thread_local std::string e; return e; }
static inline cudaError_t cudaGetLastError() {
    if (emu::last_launch_error().empty()) return cudaSuccess;
    emu::reported_error() = emu::last_launch_error();
    emu::last_launch_error().clear();
    return cudaErrorLaunchFailure;
}
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof *p);
    strcpy(p->name, "emulated B200 (tests/emu)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->l2CacheSize = 126 << 20;
    p->persistingL2CacheMaxSize = 0;  // no L2 persistence in the emulation: that branch is configuration only
    p->accessPolicyMaxWindowSize = 0; p->sharedMemPerBlockOptin = 227 * 1024; p->totalGlobalMem = (size_t)180 << 30;
    return cudaSuccess;
}
static inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
#ifdef __SANITIZE_ADDRESS__
    *p = (T*)malloc(bytes);  // exact size: AddressSanitizer then flags the first byte read or written past a "device" buffer
#else
    *p = (T*)aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
#endif
    if (!*p) return cudaErrorMemoryAllocation;
    memset((void*)*p, 0xCD, bytes);  // poison
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }

static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaGetLastError(); }
static inline cudaError_t cudaDeviceSynchronize() { return cudaGetLastError(); }
typedef void* cudaEvent_t;   /* streams are synchronous here: events are always complete */
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)malloc(8); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaStreamSetAttribute(cudaStream_t, cudaStreamAttrID, const cudaStreamAttrValue*) { return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    a->type = cudaMemoryTypeDevice; a->device = 0; a->devicePointer = (void*)p; a->hostPointer = nullptr;
    return cudaSuccess;
}
static inline cudaError_t cudaHostRegister(void* p, size_t n, unsigned) { std::lock_guard<std::mutex> g(emu::reg_mu()); emu::registered()[p] = n; return cudaSuccess; }
static inline cudaError_t cudaHostUnregister(void* p) { std::lock_guard<std::mutex> g(emu::reg_mu()); return emu::registered().erase(p) ? cudaSuccess : cudaErrorInvalidValue; }
// pinned allocations: mapped ones are registered so that cudaHostGetDevicePointer finds them
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned flags) {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess && (flags & cudaHostAllocMapped)) { std::lock_guard<std::mutex> g(emu::reg_mu()); emu::registered()[(void*)*p] = bytes; }
    return e;
}
static inline cudaError_t cudaFreeHost(void* p) { { std::lock_guard<std::mutex> g(emu::reg_mu()); emu::registered().erase(p); } free(p); return cudaSuccess; }
// zero-copy is available only inside registered ranges (as with cudaHostRegisterMapped)
template <class T> static inline cudaError_t cudaHostGetDevicePointer(T** dev, void* host, unsigned) {
    std::lock_guard<std::mutex> g(emu::reg_mu());
    for (auto& kv : emu::registered())
        if ((char*)host >= (char*)kv.first && (char*)host < (char*)kv.first + kv.second) { *dev = (T*)host; return cudaSuccess; }
    return cudaErrorInvalidValue;
}
