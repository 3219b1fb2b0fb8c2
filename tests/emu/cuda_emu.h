// cuda_emu.h -- TEST-ONLY: a minimal functional emulation of the CUDA execution model and of the
// slice of the CUDA runtime that simfire_b200/csrc/sfb.cu uses, so that the *same* kernel and
// host sources can be compiled with g++ and exercised by the CPU test-suite (no GPU in the dev
// container).  It is never built into, shipped with or loaded by the product library: the
// product has no CPU path (tests/test_cabi_cpu.py::test_no_cpu_fallback).  See tests/emu/README.md.
//
// Execution model: a launch runs its blocks one after the other; the threads of a block are
// fibers on one OS thread, switched only inside warp / block collectives (__shfl*_sync,
// __ballot_sync, __any_sync, __all_sync, __syncwarp, __syncthreads).  A collective completes when
// every live lane of the warp has arrived, so warp-uniform control flow behaves as on the device
// and a divergent collective dead-locks loudly (the scheduler aborts).  Atomics are plain
// read-modify-writes; streams, events and copies are synchronous; a captured "graph" is the list
// of recorded launches.  TMA boxes are copied synchronously at issue (zero-filled outside the
// tensor, like the hardware), mbarrier waits are no-ops.  This checks logic, not timing or races.
#pragma once
#define SFB_EMU 1

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

// ---------------------------------------------------------------------------------------
// language surface
// ---------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static  // blocks run one at a time
#define __grid_constant__

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct __attribute__((aligned(16))) uint4 {
    uint32_t x, y, z, w;
};
struct __attribute__((aligned(16))) float4 {
    float x, y, z, w;
};
struct __attribute__((aligned(16))) int4 {
    int x, y, z, w;
};
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

template <class A, class B>
static inline typename std::common_type<A, B>::type min(A a, B b) {
    using T = typename std::common_type<A, B>::type;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <class A, class B>
static inline typename std::common_type<A, B>::type max(A a, B b) {
    using T = typename std::common_type<A, B>::type;
    return (T)a > (T)b ? (T)a : (T)b;
}

static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(uint32_t v) { return __builtin_ffs((int)v); }
template <class T>
static inline T __ldg(const T* p) { return *p; }
static inline void __threadfence() {}
static inline void __threadfence_system() {}
template <class T, class V>
static inline T atomicAdd(T* p, V v) {
    T old = *p;
    *p = (T)(old + (T)v);
    return old;
}

template <class T, class V>
static inline T atomicOr(T* p, V v) {
    T old = *p;
    *p = (T)(old | (T)v);
    return old;
}
template <class T, class V>
static inline T atomicAnd(T* p, V v) {
    T old = *p;
    *p = (T)(old & (T)v);
    return old;
}
static inline long long __double_as_longlong(double d) {
    long long v;
    memcpy(&v, &d, 8);
    return v;
}
static inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, 8);
    return d;
}

// ---------------------------------------------------------------------------------------
// fibers + warps
// ---------------------------------------------------------------------------------------
namespace emu {

struct Warp {
    int live = 0, arrived = 0;
    unsigned gen = 0;
    uint32_t alive_mask = 0;
    uint64_t slot[2][32];
    uint32_t vote[2] = {0, 0};  // ballot of the lanes that took part, fixed when the collective completes
    int tag[2] = {0, 0};
};

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    uint3 tid{0, 0, 0};
    int lane = 0;
    Warp* warp = nullptr;
    bool done = false;
};

struct Block {
    int live = 0, arrived = 0;
    unsigned gen = 0;
};

struct State {
    Fiber* cur = nullptr;
    void* sched_sp = nullptr;
    uint3 block_idx{0, 0, 0};
    dim3 block_dim, grid_dim;
    Block block;
    const std::function<void()>* body = nullptr;
    unsigned char* dyn_smem = nullptr;
    size_t dyn_smem_bytes = 0;
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    long long launches = 0, switches = 0;
    // graph capture
    bool capturing = false;
    std::vector<std::function<void()>>* capture_into = nullptr;
};
State& st();

void yield();
void warp_sync(int tag);
void block_sync();
uint64_t exchange(uint64_t v, int src_lane, int tag);
uint32_t ballot(bool pred, int tag);
void run_grid(const char* name, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);

// kernel launch: arguments are evaluated and copied NOW (as cudaLaunchKernel does), the call is
// executed now or, during stream capture, recorded
template <class... P, class... A>
static inline void launch(const char* name, dim3 grid, dim3 block, size_t smem, void (*kernel)(P...), A&&... a) {
    std::tuple<typename std::decay<P>::type...> args(std::forward<A>(a)...);
    std::function<void()> body = [kernel, args]() { std::apply(kernel, args); };
    State& s = st();
    if (s.capturing) {
        s.capture_into->push_back([name, grid, block, smem, body]() { run_grid(name, grid, block, smem, body); });
        return;
    }
    run_grid(name, grid, block, smem, body);
}

// ---- TMA / mbarrier stand-ins (see header comment) ----
struct TensorMap {
    void* base;
    uint64_t dims[3];     // elements (4 bytes each), innermost first
    uint64_t strides[2];  // bytes, dims 1 and 2
    uint32_t box[3];
};
static inline uint32_t smem_u32(const void* p) {
    return (uint32_t)((const unsigned char*)p - st().dyn_smem) + 4096u;
}
static inline void* smem_ptr(uint32_t a) { return st().dyn_smem + (a - 4096u); }
static inline void mbar_init(uint32_t, uint32_t) {}
static inline void mbar_init_fence() {}
static inline void mbar_expect_tx(uint32_t, uint32_t) {}
// the copy was made when the elected lane issued it; the wait only has to order the other
// lanes behind that lane (on the device the mbarrier does)
static inline void mbar_wait(uint32_t, uint32_t) { warp_sync(8); }
static inline void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t) {
    const TensorMap& t = *reinterpret_cast<const TensorMap*>(tmap);
    uint32_t* out = reinterpret_cast<uint32_t*>(smem_ptr(dst));
    for (uint32_t k = 0; k < t.box[2]; ++k)
        for (uint32_t r = 0; r < t.box[1]; ++r)
            for (uint32_t e = 0; e < t.box[0]; ++e) {
                const long long x = (long long)c0 + e, y = (long long)c1 + r, z = (long long)c2 + k;
                uint32_t v = 0;
                if (x >= 0 && x < (long long)t.dims[0] && y >= 0 && y < (long long)t.dims[1] && z >= 0 && z < (long long)t.dims[2])
                    memcpy(&v, (const char*)t.base + z * t.strides[1] + y * t.strides[0] + x * 4, 4);
                *out++ = v;
            }
}

}  // namespace emu

#define threadIdx (emu::st().cur->tid)
#define blockIdx (emu::st().block_idx)
#define blockDim (emu::st().block_dim)
#define gridDim (emu::st().grid_dim)

static inline void __syncwarp(uint32_t = 0xffffffffu) { emu::warp_sync(1); }
static inline void __syncthreads() { emu::block_sync(); }
static inline uint32_t __ballot_sync(uint32_t, bool pred) { return emu::ballot(pred, 2); }
static inline int __any_sync(uint32_t, bool pred) { return emu::ballot(pred, 3) != 0; }
static inline int __all_sync(uint32_t, bool pred) { return emu::ballot(!pred, 4) == 0; }
static inline uint32_t __reduce_or_sync(uint32_t, uint32_t v) {
    uint32_t r = 0;
    for (int b = 0; b < 32; ++b) r |= (emu::ballot((v >> b) & 1u, 9 + b) != 0 ? 1u : 0u) << b;  // 32 votes: slow, simple
    return r;
}
static inline unsigned int __reduce_add_sync(uint32_t, unsigned int v) {
    unsigned int r = 0;
    for (int b = 0; b < 32; ++b) r += (unsigned int)__builtin_popcount(emu::ballot((v >> b) & 1u, 50 + b)) << b;  // bit-sliced sum
    return r;
}
// (width: a power of two; lanes exchange within their aligned group of `width` lanes, as the hardware does)
template <class T>
static inline T __shfl_sync(uint32_t, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
    const int lane = emu::st().cur->lane, base = lane & ~(width - 1);
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = emu::exchange(raw, base + (src & (width - 1)), 5);
    memcpy(&v, &raw, sizeof(T));
    return v;
}
template <class T>
static inline T __shfl_down_sync(uint32_t, T v, unsigned delta, int width = 32) {
    const int lane = emu::st().cur->lane, base = lane & ~(width - 1);
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = emu::exchange(raw, lane + (int)delta < base + width ? lane + (int)delta : lane, 6);
    memcpy(&v, &raw, sizeof(T));
    return v;
}
template <class T>
static inline T __shfl_up_sync(uint32_t, T v, unsigned delta, int width = 32) {
    const int lane = emu::st().cur->lane, base = lane & ~(width - 1);
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    raw = emu::exchange(raw, lane - (int)delta >= base ? lane - (int)delta : lane, 7);
    memcpy(&v, &raw, sizeof(T));
    return v;
}

// ---------------------------------------------------------------------------------------
// runtime surface (everything is synchronous and lives in host memory)
// ---------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorNotSupported = 801 };
typedef struct emuStream* cudaStream_t;
typedef struct emuEvent* cudaEvent_t;
typedef std::vector<std::function<void()>>* cudaGraph_t;
typedef std::vector<std::function<void()>>* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2, cudaHostAllocPortable = 1 };
enum { cudaStreamCaptureModeThreadLocal = 1, cudaEnableDefault = 0, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
struct cudaDeviceProp {
    int multiProcessorCount;
};
struct cudaIpcMemHandle_t {
    char reserved[64];
};

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
static inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = 1;
    return cudaSuccess;
}
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    const char* e = getenv("SFB_EMU_SMS");
    p->multiProcessorCount = e ? std::max(1, atoi(e)) : 2;
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t bytes) {
#if defined(__SANITIZE_ADDRESS__)
    // exact size, so that AddressSanitizer's red zone starts at the first byte past the allocation
    if (posix_memalign(p, 256, bytes ? bytes : 1) != 0) *p = nullptr;
#else
    *p = aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
#endif
    if (*p) memset(*p, 0xA5, bytes);  // device memory is not zero-initialised
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
static inline cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) {
    *free_b = (size_t)8 << 30;
    *total_b = (size_t)16 << 30;
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) {
    free(p);
    return cudaSuccess;
}
static inline cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
static inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
static inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) {
    *d = h;
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    memmove(d, s, n);
    return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
    memset(d, v, n);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = reinterpret_cast<cudaStream_t>(malloc(8));
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    free(s);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) {
    *lo = 0;
    *hi = -5;
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = reinterpret_cast<cudaEvent_t>(malloc(8));
    return cudaSuccess;
}
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    free(e);
    return cudaSuccess;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) {
    *ms = 0.001f;
    return cudaSuccess;
}
template <class K>
static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
template <class K>
static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, K, int, size_t) {
    *n = 2;
    return cudaSuccess;
}
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) {
    emu::State& s = emu::st();
    s.capturing = true;
    s.capture_into = new std::vector<std::function<void()>>();
    return cudaSuccess;
}
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
    emu::State& s = emu::st();
    *g = s.capture_into;
    s.capturing = false;
    s.capture_into = nullptr;
    return cudaSuccess;
}
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) {
    *e = new std::vector<std::function<void()>>(*g);
    return cudaSuccess;
}
static inline cudaError_t cudaGraphDestroy(cudaGraph_t g) {
    delete g;
    return cudaSuccess;
}
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) {
    delete e;
    return cudaSuccess;
}
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) {
    for (auto& f : *e) f();
    return cudaSuccess;
}
// "IPC" inside one process: the handle carries the pointer
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    memset(h, 0, sizeof(*h));
    memcpy(h->reserved, &p, sizeof(p));
    return cudaSuccess;
}
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    memcpy(p, h.reserved, sizeof(*p));
    return cudaSuccess;
}
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// ---- driver API slice: tensor maps ----
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
typedef uint64_t cuuint64_t;
typedef uint32_t cuuint32_t;
struct __attribute__((aligned(64))) CUtensorMap {
    emu::TensorMap t;
    char pad[128 - sizeof(emu::TensorMap)];
};
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_UINT32 = 4 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_128B = 2 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
static inline CUresult emuTensorMapEncodeTiled(CUtensorMap* m, CUtensorMapDataType, cuuint32_t rank, void* base,
                                               const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box,
                                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (rank != 3) return 1;
    memset(m, 0, sizeof(*m));
    m->t.base = base;
    for (int i = 0; i < 3; ++i) {
        m->t.dims[i] = dims[i];
        m->t.box[i] = box[i];
    }
    m->t.strides[0] = strides[0];
    m->t.strides[1] = strides[1];
    return CUDA_SUCCESS;
}
static inline cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long,
                                                  cudaDriverEntryPointQueryResult* q) {
    *fn = strcmp(name, "cuTensorMapEncodeTiled") == 0 ? reinterpret_cast<void*>(&emuTensorMapEncodeTiled) : nullptr;
    if (q) *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------
// implementation (one translation unit defines CUDA_EMU_IMPLEMENTATION)
// ---------------------------------------------------------------------------------------
#ifdef CUDA_EMU_IMPLEMENTATION
#if !defined(__x86_64__)
#error "the fiber switch below is written for x86-64"
#endif
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

static State g_state;
State& st() { return g_state; }

static constexpr size_t STACK_BYTES = 256 << 10;

[[noreturn]] static void die(const char* what) {
    fprintf(stderr, "cuda_emu: %s\n", what);
    abort();
}

static void release_if_complete(Warp* w) {
    if (w->arrived > 0 && w->arrived == w->live) {
        const int par = w->gen & 1;
        uint32_t m = 0;
        for (int l = 0; l < 32; ++l)
            if (((w->alive_mask >> l) & 1u) && (w->slot[par][l] & 1u)) m |= 1u << l;
        w->vote[par] = m;
        w->arrived = 0;
        w->gen++;
    }
}

static void fiber_exit() {
    State& s = g_state;
    Fiber* f = s.cur;
    f->done = true;
    static const bool trace = getenv("SFB_EMU_TRACE") != nullptr && atoi(getenv("SFB_EMU_TRACE")) >= 3;
    if (trace) fprintf(stderr, "[emu] exit block %u thread %u (warp gen %u)\n", s.block_idx.x, f->tid.x, f->warp->gen);
    Warp* w = f->warp;
    w->live--;
    w->alive_mask &= ~(1u << f->lane);
    release_if_complete(w);
    s.block.live--;
    if (s.block.arrived > 0 && s.block.arrived == s.block.live) {
        s.block.arrived = 0;
        s.block.gen++;
    }
    emu_switch(&f->sp, s.sched_sp);
    die("a finished fiber was resumed");
}

extern "C" void emu_fiber_entry() {
    (*g_state.body)();
    fiber_exit();
}

void yield() {
    State& s = g_state;
    s.switches++;
    static const bool trace = getenv("SFB_EMU_TRACE") != nullptr && atoi(getenv("SFB_EMU_TRACE")) >= 2;
    if (trace && (s.switches % 2000000) == 0) {
        Warp* w = s.cur->warp;
        fprintf(stderr, "[emu] %lld switches; block %u thread %u parked at collective tag %d (warp gen %u, arrived %d of %d live)\n",
                s.switches, s.block_idx.x, s.cur->tid.x, w->tag[w->gen & 1], w->gen, w->arrived, w->live);
    }
    emu_switch(&s.cur->sp, s.sched_sp);
}

void warp_sync(int tag) {
    Warp* w = g_state.cur->warp;
    const unsigned g = w->gen;
    const int par = g & 1;
    if (w->arrived == 0) w->tag[par] = tag;
    else if (w->tag[par] != tag) {
        fprintf(stderr, "cuda_emu: block %u thread %u arrives at collective tag %d, its warp is parked at tag %d\n",
                g_state.block_idx.x, g_state.cur->tid.x, tag, w->tag[par]);
        die("lanes of one warp met at different collectives (divergent collective)");
    }
    w->arrived++;
    release_if_complete(w);
    while (w->gen == g) yield();
}

void block_sync() {
    Block& b = g_state.block;
    const unsigned g = b.gen;
    if (++b.arrived == b.live) {
        b.arrived = 0;
        b.gen++;
        return;
    }
    while (b.gen == g) yield();
}

uint64_t exchange(uint64_t v, int src_lane, int tag) {
    Fiber* f = g_state.cur;
    Warp* w = f->warp;
    const int par = w->gen & 1;
    w->slot[par][f->lane] = v;
    warp_sync(tag);
    // (a lane may leave the kernel right after a collective: its slot stays readable)
    return w->slot[par][src_lane];
}

uint32_t ballot(bool pred, int tag) {
    Fiber* f = g_state.cur;
    Warp* w = f->warp;
    const int par = w->gen & 1;
    w->slot[par][f->lane] = pred ? 1 : 0;
    warp_sync(tag);
    return w->vote[par];
}

void run_grid(const char* name, dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
    State& s = g_state;
    static const bool trace = getenv("SFB_EMU_TRACE") != nullptr;
    if (trace) fprintf(stderr, "[emu] %s <<<%u, %u, %zu>>>\n", name, grid.x, block.x, smem);
    if (s.cur) die("nested launch");
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads < 1 || nthreads > 1024) die("block size");
    if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) die("only 1-D launches are emulated");
    s.launches++;
    if ((int)s.fibers.size() < nthreads) {
        const size_t old = s.fibers.size();
        s.fibers.resize(nthreads);
        for (size_t i = old; i < s.fibers.size(); ++i) {
            s.fibers[i].stack = (char*)aligned_alloc(4096, STACK_BYTES);
            if (!s.fibers[i].stack) die("out of memory (fiber stacks)");
        }
    }
    const int nwarps = (nthreads + 31) / 32;
    s.warps.resize(nwarps);
    if (s.dyn_smem_bytes < smem + 4096) {
        free(s.dyn_smem);
        s.dyn_smem_bytes = smem + 4096;
        s.dyn_smem = (unsigned char*)aligned_alloc(4096, (s.dyn_smem_bytes + 4095) / 4096 * 4096);
    }
    s.block_dim = block;
    s.grid_dim = grid;
    s.body = &body;
    for (unsigned b = 0; b < grid.x; ++b) {
        s.block_idx = uint3{b, 0, 0};
        s.block = Block();
        s.block.live = nthreads;
        memset(s.dyn_smem, 0xCD, s.dyn_smem_bytes);  // shared memory is not initialised
        for (int w = 0; w < nwarps; ++w) {
            Warp& W = s.warps[w];
            W = Warp();
            W.live = std::min(32, nthreads - 32 * w);
            W.alive_mask = W.live == 32 ? 0xffffffffu : ((1u << W.live) - 1u);
        }
        for (int t = 0; t < nthreads; ++t) {
            Fiber& f = s.fibers[t];
            f.tid = uint3{(unsigned)t, 0, 0};
            f.lane = t & 31;
            f.warp = &s.warps[t >> 5];
            f.done = false;
            // initial frame for emu_switch: six callee-saved registers, then the entry point as the
            // return address; at the entry rsp % 16 == 8 as after a call
            uintptr_t top = ((uintptr_t)f.stack + STACK_BYTES) & ~(uintptr_t)15;
            void** sp = reinterpret_cast<void**>(top);
            *--sp = nullptr;                                        // fake return address of the entry
            *--sp = reinterpret_cast<void*>(&emu_fiber_entry);      // `ret` target
            for (int r = 0; r < 6; ++r) *--sp = nullptr;
            f.sp = sp;
        }
        int remaining = nthreads;
        auto signature = [&]() {
            unsigned long long sig = (unsigned long long)s.block.gen * 1031ull + (unsigned long long)s.block.arrived;
            for (int w = 0; w < nwarps; ++w) sig = sig * 1000003ull + (unsigned long long)s.warps[w].gen * 64ull + (unsigned long long)s.warps[w].arrived;
            return sig * 4099ull + (unsigned long long)remaining;
        };
        while (remaining > 0) {
            const unsigned long long before = signature();
            for (int t = 0; t < nthreads; ++t) {
                Fiber& f = s.fibers[t];
                if (f.done) continue;
                s.cur = &f;
                emu_switch(&s.sched_sp, f.sp);
                s.cur = nullptr;
                if (f.done) --remaining;
            }
            // a whole round in which no fiber finished and no collective moved: every live fiber is
            // parked in a collective that can never complete
            if (remaining > 0 && signature() == before) die("dead-lock: a collective was not reached by every live lane");
        }
    }
    s.body = nullptr;
}

}  // namespace emu
#endif  // CUDA_EMU_IMPLEMENTATION
