// libsimfire_b200.so -- host side of the C ABI declared in include/simfire_b200.h.
// Plain CUDA runtime; no torch, no CPU fallback: every entry point that computes launches
// kernels on the handle's device and fails with SFB_ERR_CUDA if it cannot.
#ifdef SFB_EMU
// test-only build with g++ (tests/emu): the fiber emulator stands in for the CUDA runtime
#include "cuda_emu.h"
#define SFB_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(#kernel, (grid), (block), (smem), kernel, __VA_ARGS__)
#define SFB_LAUNCH_DEP(dep, kernel, grid, block, smem, stream, ...) emu::launch(#kernel, (grid), (block), (smem), kernel, __VA_ARGS__)
#else
#include <cuda_runtime.h>
#define SFB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream drains; it
// executes griddepcontrol.wait (grid_dep_wait) before it touches anything the predecessor wrote
template <typename... KArgs, typename... Args>
static inline void sfb_launch_dep(bool dep, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = dep ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#define SFB_LAUNCH_DEP(dep, kernel, grid, block, smem, stream, ...) sfb_launch_dep((dep), kernel, (grid), (block), (smem), (stream), __VA_ARGS__)
#endif
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/simfire_b200.h"
#include "sfb_kernels.cuh"
#include "sfb_lists.cuh"
#include "sfb_bits.cuh"
#if !defined(SFB_EMU) && defined(SFB_EXPERIMENT_SORT)
#include <cub/device/device_radix_sort.cuh>
#endif

using namespace sfb;

// ---------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(_e == cudaErrorMemoryAllocation ? SFB_ERR_NOMEM : SFB_ERR_CUDA,           \
                        "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ---------------------------------------------------------------------------------------
// host worker pool (patching the host mirror from the change log)
// ---------------------------------------------------------------------------------------
// Workers spin for a short while after a job before they go to sleep on the condition variable:
// sfb_sync_fire_maps issues one job per env group per step, a fraction of a millisecond apart,
// and a futex wake-up per job would cost more than the patching itself.
class HostPool {
  public:
    explicit HostPool(unsigned n) : n_(std::max(1u, n)) {
        for (unsigned t = 1; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> g(m_);
            stop_.store(true, std::memory_order_relaxed);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        for (auto& w : workers_) w.join();
    }
    unsigned size() const { return n_; }
    // runs fn(t) for t in [0, size()) and returns when all are done
    void run(const std::function<void(unsigned)>& fn) {
        fn_ = &fn;
        pending_.store(n_ - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> g(m_);  // a worker about to sleep re-checks gen_ under this mutex
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        fn(0);
        for (unsigned spins = 0; pending_.load(std::memory_order_acquire) != 0;) relax(++spins);
        fn_ = nullptr;
    }
    // all size() threads of the running job meet here (call it the same number of times in each)
    void barrier() {
        const unsigned g = bar_gen_.load(std::memory_order_acquire);
        if (bar_count_.fetch_add(1, std::memory_order_acq_rel) + 1 == n_) {
            bar_count_.store(0, std::memory_order_relaxed);
            bar_gen_.fetch_add(1, std::memory_order_release);
        } else {
            for (unsigned spins = 0; bar_gen_.load(std::memory_order_acquire) == g;) relax(++spins);
        }
    }

  private:
    static void relax(unsigned spins) {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0xFFFu) == 0) std::this_thread::yield();  // more threads than cores: let the others run
    }
    void loop(unsigned t) {
        using clock = std::chrono::steady_clock;
        unsigned long long seen = 0;
        for (;;) {
            const auto idle_since = clock::now();
            for (unsigned spins = 1; gen_.load(std::memory_order_acquire) == seen; ++spins) {
                relax(spins);
                if ((spins & 0x3FFu) == 0 && clock::now() - idle_since > std::chrono::microseconds(SPIN_US)) {
                    std::unique_lock<std::mutex> l(m_);
                    cv_.wait(l, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                    break;
                }
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_.load(std::memory_order_relaxed)) return;
            (*fn_)(t);
            pending_.fetch_sub(1, std::memory_order_release);
        }
    }
    static constexpr int SPIN_US = 1000;
    unsigned n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_;
    const std::function<void(unsigned)>* fn_ = nullptr;
    std::atomic<unsigned> pending_{0}, bar_count_{0}, bar_gen_{0};
    std::atomic<unsigned long long> gen_{0};
    std::atomic<bool> stop_{false};
};

// ---------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------
// Envs never interact, so a handle steps them as G independent groups on G streams: while one
// group is in its issue-bound kernels (k_rows, k_eval) another one streams (k_sweep), and the
// hardware overlaps the two.  Each group sees its envs through a DevParams view with offset
// plane pointers and its own work queue, row-task list and counters.
constexpr int N_COUNTERS = 12;
struct EnvGroup {
    DevParams d;
    CUtensorMap tmap;
    cudaStream_t stream;       // equal priority: best overlap between the groups
    cudaStream_t stream_prio;  // descending priority: groups finish one after the other (change log on)
    cudaStream_t last_stream;  // the one the last steps ran on
    cudaEvent_t done;
    unsigned long long* counters;  // qcount[2] | unit_next[2] | rows_count[2] | rows_next[2] | overflow | - | units_count[2]
    int sweep_blocks, rows_blocks, units_blocks;
};

struct sfb_sim {
    sfb_params prm;
    DevParams d;   // the whole handle (setup / conversion kernels, group 0's queues)
    std::vector<EnvGroup> groups;  // G >= 1 views of disjoint env ranges (slices of the queue / row list)
    EnvGroup all;                  // one view of every env (whole queue / row list); used when the
                                   // groups would not overlap anyway (kernel timing, change log on)
    int last_mode;                 // 0 none yet, 1 `all`, 2 `groups`
    // two consecutive steps (parity 0 then 1) of the whole-handle view as a CUDA graph: small
    // batches are launch-bound, and a graph replay costs less than six kernel launches
    cudaGraphExec_t pair_graph;
    unsigned pair_graph_epoch, view_epoch;  // the graph bakes the view's parameters in
    // the same for multi-group handles (opt-in, SFB_STEP_GRAPH): two steps of every group, forked over the
    // group streams (only while the change log is off; with it on the host waits for the groups one by one)
    cudaGraphExec_t group_graph;
    unsigned group_graph_epoch;
    int group_graph_on;
    int group_graph_steps;   // steps per replay (even: the graph starts and ends at parity 0)
    int64_t group_launches_all, group_launches_step;
    int64_t pair_launches_all, pair_launches_step;
    cudaEvent_t fork_ev;
    int cell_bytes;   // 1 or 2
    int use_tma;      // sweep front end
    int sweep_blocks; // persistent grid of the sweep kernel
    int rows_blocks;  // persistent grid of k_rows
    int unit_skip;    // the sweep only reads flagged units (DevParams::unit_act)
    int unit_rows;    // ... and a unit is a single row of a strip: no sweep at all (k_row_list)
    int pdl;          // bitboard handles: programmatic dependent launch of the step kernels (SFB_PDL=0 turns it off)
    int front_bits;   // bitboard step (sfb_bits.cuh): k_tiles + k_eval instead of k_row_list + k_rows + k_eval
    int tiles_blocks; // grid of k_tiles
    void (*tiles_fn)(DevParams, int, int);  // k_tiles<cell type, max_fire_duration, no statistics>
    void (*tiles_fn_stats)(DevParams, int, int);
    int ext_writes;   // bitboard handles: a status was written from outside (mitigation, map upload) since all envs
                      // were last reset: control lines may exist, so a step needs k_eval; without, k_tiles closes the step
    int fuse_eval;    // SFB_FUSE_EVAL=1 turns the one-kernel step on (measured slower than two kernels: off by default)
    int front_lists;  // list-driven step (sfb_lists.cuh): k_front (+ k_tail with attenuation); no env groups
    int lpar;         // which watch-list buffer the NEXT step reads
    int front_blocks; // persistent grid of k_front
    uint8_t* env_mark;            // device [E]: envs a purge / rebuild of the watch list applies to
    unsigned long long* list_ctr; // device: wl_count[2] | ros_count | broken (int32) + ticket (uint32)
    int64_t list_entries_last;    // entries the last completed step left on the list (sfb_get_queue_stats)
    CUtensorMap tmap; // state plane as uint32 [E][H][pitch_bytes / 4]
    int parity;       // which half of meta / qcount the NEXT step reads
    int in_step;      // sfb_step_sweep done, sfb_step_eval pending
    int static_dirty; // raw static planes changed since k_derive_static last ran
    size_t mailbox_off;   // byte offset of the slab mailbox behind the state plane
    uint32_t slab_step;   // global step counter of sfb_step_slab (same on every slab)
    int n_sm;
    cudaStream_t stream;      // stream in use
    cudaStream_t own_stream;  // created by the handle
    cudaEvent_t ev[4];
    cudaEvent_t span[2];      // sfb_step_timed
    // scratch
    void* stage;          // device staging for host <-> device plane traffic
    size_t stage_bytes;
    void* obs;            // int8 [E][H][W] observation buffer (sfb_fire_map_device)
    int32_t* small;       // device scratch for env lists / points
    size_t small_bytes;
    int32_t* pts_host[2];    // pinned staging for sfb_apply_points: the caller's buffer is free again on return
    cudaEvent_t pts_ev[2];   // ... without waiting for the device (the slot's previous copy is waited for instead)
    int pts_slot;
    // accounting
    int64_t launches_all, launches_step;
    int64_t dev_bytes;
    int timing;
    double sweep_ms, rows_ms, eval_ms;
    int64_t timed_steps;
    int64_t last_entries;
    int last_overflow;
    // host mirror bookkeeping (sfb_sync_fire_maps)
    long long patch_parallel_min;  // logs shorter than this are patched by the calling thread alone
    int steps_since_sync;      // update() calls enqueued since the change logs were last drained
    int setup_after_step;      // ... and a reset / mitigation call came after one of them
    const int8_t* mirror;      // buffer the last sync wrote, nullptr = none valid
    int full_resync;           // something changed that the log does not describe
    unsigned long long* log_head;    // pinned: {count, overflow} of every log, read back each sync
    int head_valid;                  // the step that ran last already copied every group's head into log_head
    unsigned long long* log_mapped[MAX_ENV_GROUPS];  // the change logs (mapped host memory)
    unsigned long long* log_counts;  // device: {count, overflow} x n_logs
    HostPool* pool;
    std::vector<std::vector<unsigned long long>> buckets;  // [chunk * T + owner]
};

static int use(sfb_sim* s) {
    CU(cudaSetDevice(s->prm.device));
    s->head_valid = 0;  // any call may append to the change logs; only a grouped step re-validates
    s->d.bits_par = s->parity;  // bitboard handles: setup kernels put tiles on the list the next step reads
    return 0;
}

template <typename T>
static int dmalloc(sfb_sim* s, T** p, size_t bytes) {
    CU(cudaMalloc((void**)p, bytes));
    s->dev_bytes += (int64_t)bytes;
    return 0;
}

static int ensure_stage(sfb_sim* s, size_t bytes) {
    if (s->stage_bytes >= bytes) return 0;
    if (s->stage) {
        CU(cudaFree(s->stage));
        s->dev_bytes -= (int64_t)s->stage_bytes;
        s->stage = nullptr;
        s->stage_bytes = 0;
    }
    int rc = dmalloc(s, &s->stage, bytes);
    if (rc) return rc;
    s->stage_bytes = bytes;
    return 0;
}

static constexpr size_t PTS_STAGE_BYTES = 64 << 10;

static int ensure_small(sfb_sim* s, size_t bytes) {
    if (s->small_bytes >= bytes) return 0;
    if (s->small) {
        CU(cudaStreamSynchronize(s->stream));
        CU(cudaFree(s->small));
        s->dev_bytes -= (int64_t)s->small_bytes;
        s->small = nullptr;
        s->small_bytes = 0;
    }
    bytes = std::max(bytes, (size_t)1 << 16);
    int rc = dmalloc(s, &s->small, bytes);
    if (rc) return rc;
    s->small_bytes = bytes;
    return 0;
}

static inline unsigned nblocks(long long n, int threads) { return (unsigned)((n + threads - 1) / threads); }
// byte cells and no row padding: a plane is one linear run of H * W bytes
static inline bool linear_bytes(const sfb_sim* s) { return s->cell_bytes == 1 && s->d.pitch == s->d.W; }

// ---------------------------------------------------------------------------------------
// setup / conversion kernels (not on the per-step path)
// ---------------------------------------------------------------------------------------
template <typename CellT>
__global__ void k_clear_envs(DevParams p, const int32_t* envs, int n, int clear_burn) {
    const long long total = (long long)n * p.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / p.plane);
        const long long cell = i - (long long)k * p.plane;
        const int env = envs ? envs[k] : k;
        const int x = (int)(cell % p.pitch);
        const long long idx = (long long)env * p.plane + cell;
        reinterpret_cast<CellT*>(p.state)[idx] = (CellT)(x < p.W ? ST_UNBURNED : ST_BURNED);
        if (clear_burn) {
            p.burn[idx] = 0.0;
            if (p.ros) p.ros[idx] = 0.0;
            if (p.ign) p.ign[idx] = -1;
        }
    }
}

// FireSimulation._create_fire_map + FireManager.__init__ (simulation.py:561-566, fire.py:101-103)
template <typename CellT>
__global__ void k_reset_meta(DevParams p, int par, const int32_t* envs, const int32_t* xy, int n, int y_off, int lpar) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int env = envs ? envs[k] : k;
    const int x = xy[2 * k], y = xy[2 * k + 1] - y_off;
    if (y >= 0 && y < p.H) {
        reinterpret_cast<CellT*>(p.state)[(long long)env * p.plane + (long long)y * p.pitch + x] =
            (CellT)(ST_BURNING | (1 << 3));  // sprite created before update() call 1: ign = 0
        if (p.ign) p.ign[(long long)env * p.plane + (long long)y * p.pitch + x] = 0;
        if (p.unit_act) mark_units_around<CellT>(p, env, y, x);
        // list handles: the sprite joins the watch list; at duration 0 its thread brings the neighbours in
        if (p.listed && listed_test_and_set(p, (long long)env * p.plane + (long long)y * p.pitch + x)) list_append(p, lpar, env, y, x);
    }
    EnvMeta m;
    m.t = 1;
    m.running = 1;
    m.elapsed = 0.0;
    m.any_live = m.any_cand = m.pad = 0;
    m.time_quit = p.has_max_time && (p.dt > p.max_time || 0.0 > p.max_time);
    p.meta[(long long)par * p.meta_stride + env] = m;
    if (p.track) {  // "env was cleared", then its first burning cell, in this order
        const LogRef& L = log_of_env(p, env);
        const unsigned long long slot = atomicAdd(L.count, 2ULL);
        const unsigned long long reset = (unsigned long long)env | ((unsigned long long)LOG_ENV_RESET << 48) | LOG_SETUP_BIT;
        const bool inside = y >= 0 && y < p.H;
        const long long idx = (long long)env * p.plane + (long long)(inside ? y : 0) * p.pitch + x;
        log_put(L, slot, reset);
        // a slab that does not hold the ignition row logs the reset twice (harmless)
        log_put(L, slot + 1, inside ? ((unsigned long long)idx | (1ULL << 48) | LOG_SETUP_BIT) : reset);
    }
}

// ControlLineManager.update (mitigation.py:77): fire_map[y, x] = kind, sprite untouched
template <typename CellT>
__global__ void k_apply_points(DevParams p, const int32_t* pts, long long n, int kind, int y_off, int lpar) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int env = pts[4 * i], x = pts[4 * i + 1], y = pts[4 * i + 2] - y_off, k = pts[4 * i + 3];
    if (k != kind || y < 0 || y >= p.H) return;
    const long long idx = (long long)env * p.plane + (long long)y * p.pitch + x;
    CellT* c = reinterpret_cast<CellT*>(p.state) + idx;
    const int nc = (*c & ~7) | to_internal(k);
    *c = (CellT)nc;
    if (p.unit_act) mark_units_around<CellT>(p, env, y, x);  // a control line is work under attenuation
    // list handles: a cell that became ignitable next to a sprite, or an attenuated control line, is watched
    if (p.listed && belongs_on_list<CellT>(p, env, y, x, nc) && listed_test_and_set(p, idx)) list_append(p, lpar, env, y, x);
    if (p.bits) bits_on_status(p, env, y, x, nc & 7);
    if (p.track) {
        const LogRef& L = log_of_env(p, env);
        log_put(L, atomicAdd(L.count, 1ULL), (unsigned long long)idx | ((unsigned long long)k << 48) | LOG_SETUP_BIT);
    }
}

template <typename CellT>
__global__ void k_set_map(DevParams p, int env0, int n, const int8_t* maps) {
    const long long hw = (long long)p.H * p.W, total = (long long)n * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / hw);
        const long long r = i - (long long)k * hw;
        const int y = (int)(r / p.W), x = (int)(r - (long long)y * p.W);
        CellT* c = reinterpret_cast<CellT*>(p.state) + (long long)(env0 + k) * p.plane + (long long)y * p.pitch + x;
        *c = (CellT)((*c & ~7) | to_internal(maps[i] & 7));
    }
}

template <typename CellT>
__global__ void k_get_map(DevParams p, int env0, int n, int8_t* maps) {
    const long long hw = (long long)p.H * p.W, total = (long long)n * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / hw);
        const long long r = i - (long long)k * hw;
        const int y = (int)(r / p.W), x = (int)(r - (long long)y * p.W);
        const CellT c = reinterpret_cast<const CellT*>(p.state)[(long long)(env0 + k) * p.plane + (long long)y * p.pitch + x];
        maps[i] = (int8_t)to_burn_status(c & 7);
    }
}

// 16 cells per thread; valid when cells are bytes and pitch == W (plane = H * W, multiple of 16)
__global__ void k_get_map_v16(const uint4* __restrict__ state, uint4* __restrict__ maps, long long n16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        uint4 v = state[i];
        auto conv = [](uint32_t w) { const uint32_t s = w & 0x07070707u; return s - ((s >> 2) & 0x01010101u); };
        maps[i] = make_uint4(conv(v.x), conv(v.y), conv(v.z), conv(v.w));
    }
}

__global__ void k_set_map_v16(uint4* __restrict__ state, const uint4* __restrict__ maps, long long n16) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
        const uint4 m = maps[i];
        uint4 v = state[i];
        auto conv = [](uint32_t old, uint32_t e) {
            e &= 0x07070707u;
            const uint32_t in = e + (((e + 0x01010101u) >> 2) & 0x01010101u);  // BurnStatus -> internal
            return (old & 0xF8F8F8F8u) | in;
        };
        state[i] = make_uint4(conv(v.x, m.x), conv(v.y, m.y), conv(v.z, m.z), conv(v.w, m.w));
    }
}

// clears whole envs: state = UNBURNED, burn = 0 (ros = 0); same validity condition as above
__global__ void k_clear_envs_v16(DevParams p, const int32_t* envs, int n, long long plane16) {
    const long long total = (long long)n * plane16;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / plane16);
        const long long w = i - (long long)k * plane16;
        const long long env = envs ? envs[k] : k;
        reinterpret_cast<uint4*>(p.state)[env * plane16 + w] = z;
        uint4* b = reinterpret_cast<uint4*>(p.burn) + (env * plane16 + w) * 8;  // 16 cells x 8 B = 8 words
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = z;
        if (p.ros) {
            uint4* r = reinterpret_cast<uint4*>(p.ros) + (env * plane16 + w) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = z;
        }
        if (p.ign) {
            const uint4 m1 = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            uint4* g = reinterpret_cast<uint4*>(p.ign) + (env * plane16 + w) * 4;  // 16 cells x 4 B
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = m1;
        }
    }
}

template <typename CellT>
__global__ void k_get_plane(DevParams p, int par, int env, int plane, void* out) {
    const long long hw = (long long)p.H * p.W;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
    const long long idx = (long long)env * p.plane + (long long)y * p.pitch + x;
    if (plane == SFB_PLANE_BURN) {
        ((double*)out)[i] = p.burn[idx];
    } else if (plane == SFB_PLANE_ROS) {
        ((double*)out)[i] = p.ros[idx];
    } else if (plane == SFB_PLANE_IGNITION) {
        ((int32_t*)out)[i] = p.ign[idx];
    } else {
        const int c = reinterpret_cast<const CellT*>(p.state)[idx];
        if (plane == SFB_PLANE_STATUS) {
            ((int8_t*)out)[i] = (int8_t)to_burn_status(c & 7);
        } else {
            // duration as the NEXT update() call will see it before pruning
            const int t = p.meta[(long long)par * p.meta_stride + env].t;
            ((int32_t*)out)[i] = (c >> 3) ? sprite_age<CellT>(c >> 3, (t - 1) % Cell<CellT>::M) : -1;
        }
    }
}

__global__ void k_set_static(DevParams p, int env_first, int n_env, int plane, const float* src) {
    const long long hw = (long long)p.H * p.W, total = (long long)n_env * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / hw);
        const long long r = i - (long long)k * hw;
        const int y = (int)(r / p.W), x = (int)(r - (long long)y * p.W);
        float* rec = (float*)(p.stat + (long long)(env_first + k) * p.plane + (long long)y * p.pitch + x);
        rec[plane] = src[r];  // the same host plane for every env of the range
    }
}

// RothermelFireManager._compute_slopes (fire.py:436-449): np.gradient (second-order central
// differences inside, first-order one-sided at the borders), all in float64
__global__ void k_slopes(DevParams p, int env_first, int n_env, const double* elev, double spacing) {
    const long long hw = (long long)p.H * p.W, total = (long long)n_env * hw;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / hw);
        const long long r = i - (long long)k * hw;
        const int y = (int)(r / p.W), x = (int)(r - (long long)y * p.W);
        auto at = [&](int yy, int xx) { return elev[(long long)yy * p.W + xx]; };
        double gy, gx;
        if (p.H == 1) gy = 0.0;
        else if (y == 0) gy = (at(1, x) - at(0, x)) / spacing;
        else if (y == p.H - 1) gy = (at(y, x) - at(y - 1, x)) / spacing;
        else gy = (at(y + 1, x) - at(y - 1, x)) / (2.0 * spacing);
        if (p.W == 1) gx = 0.0;
        else if (x == 0) gx = (at(y, 1) - at(y, 0)) / spacing;
        else if (x == p.W - 1) gx = (at(y, x) - at(y, x - 1)) / spacing;
        else gx = (at(y, x + 1) - at(y, x - 1)) / (2.0 * spacing);
        float* rec = (float*)(p.stat + (long long)(env_first + k) * p.plane + (long long)y * p.pitch + x);
        rec[SFB_SLOPE_MAG] = (float)sqrt(gx * gx + gy * gy);
        rec[SFB_SLOPE_DIR] = (float)atan2(gy, gx + 0.000001);
    }
}

// the dense rate_of_spread plane of the reference is rebuilt from zeros each step that gets
// past the early return (fire.py:703-708)
__global__ void k_clear_ros(DevParams p, int par) {
    const long long total = (long long)p.E * p.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int env = (int)(i / p.plane);
        const EnvMeta& m = p.meta[(long long)par * p.meta_stride + env];
        if (m.running && !m.time_quit && m.any_cand) p.ros[i] = 0.0;
    }
}

// bitboard handles evaluate their candidates before the env-wide "any candidate" is known (k_tiles): the
// step's rates are written to a second plane, zeroed before k_tiles and committed after k_eval
__global__ void k_ros_zero(DevParams p) {
    const long long total = (long long)p.E * p.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        p.ros_w[i] = 0.0;
}
__global__ void k_ros_commit(DevParams p, int par) {
    const long long total = (long long)p.E * p.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const EnvMeta& m = p.meta[(long long)par * p.meta_stride + (int)(i / p.plane)];
        if (m.running && !m.time_quit && m.any_cand) p.ros[i] = p.ros_w[i];
    }
}

// fuel-only Rothermel terms of every static cell, through the same code as the one-shot path
__global__ void k_derive_static(DevParams p, long long n_cells) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += (long long)gridDim.x * blockDim.x) {
        const StaticRec r = p.stat[i];
        DerivedRec d;
        d.fuel = sfb_fuel_terms(r.fuel.x, r.fuel.y, r.fuel.z, r.fuel.w, p.part);
        d.env = r.env;
        if (p.drv) const_cast<DerivedRec*>(p.drv)[i] = d;
        if (p.rtab) {  // every input of rothermel.py:4-136 is static per (cell, direction): evaluate once
            double* out = const_cast<double*>(p.rtab) + i * 8;
#pragma unroll
            for (int dir = 0; dir < 8; ++dir) out[dir] = sfb_spread_from_terms(dir, d.fuel, d.env.x, d.env.y, d.env.z, d.env.w);
        }
    }
}

// test / measurement knob (sfb_debug_stall): keeps the handle's stream busy for a while, so that a
// missing ordering between it and another stream shows up as a wrong result instead of passing by luck
__global__ void k_stall(long long ns) {
#ifndef SFB_EMU
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while ((long long)(t - t0) < ns);
#else
    (void)ns;
#endif
}

__global__ void k_rate_of_spread(const int8_t* dir, const float* rec, SfbParticle fp, long long n, double* out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sfb_rate_of_spread_pair(dir[i], rec + 8 * i, fp);
}

// ---------------------------------------------------------------------------------------
// ConstantSpreadFireManager.update (fire.py:754-787), literally: prune (fire.py:116-161), then every sprite
// whose duration equals rate_of_spread sets its in-bounds ignitable neighbours (fire.py:163-234) to BURNING.
// The sprites it appends for them carry no duration entry and are sliced off by the next call's prune
// (fire.py:148-155), so those cells stay BURNING and never spread: on the device they get the status
// without a sprite code.  Two kernels, because the reference prunes every sprite before any spreads.
// ---------------------------------------------------------------------------------------
template <typename CellT>
__global__ void k_cs_prune(DevParams p, int par) {
    const long long total = (long long)p.E * p.plane;
    CellT* state = reinterpret_cast<CellT*>(p.state);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = state[i];
        if ((c >> 3) == 0) continue;
        const EnvMeta& m = p.meta[(long long)par * p.meta_stride + (int)(i / p.plane)];
        if (!m.running) continue;
        if (sprite_age<CellT>(c >> 3, (m.t - 1) % Cell<CellT>::M) >= p.max_dur) state[i] = (CellT)ST_BURNED;
    }
}
template <typename CellT>
__global__ void k_cs_spread(DevParams p, int par, int rate_of_spread) {
    const long long total = (long long)p.E * p.plane;
    CellT* state = reinterpret_cast<CellT*>(p.state);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = state[i];
        if ((c >> 3) == 0) continue;
        const int env = (int)(i / p.plane);
        const EnvMeta& m = p.meta[(long long)par * p.meta_stride + env];
        if (!m.running || sprite_age<CellT>(c >> 3, (m.t - 1) % Cell<CellT>::M) != rate_of_spread) continue;
        const long long cell = i - (long long)env * p.plane;
        const int y = (int)(cell / p.pitch), x = (int)(cell - (long long)y * p.pitch);
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) {
                if ((dy == 0 && dx == 0) || (!p.diagonal && dy != 0 && dx != 0)) continue;
                const int yy = y + dy, xx = x + dx;
                if ((unsigned)yy >= (unsigned)p.H || (unsigned)xx >= (unsigned)p.W) continue;
                CellT* q = state + i + (long long)dy * p.pitch + dx;
                const int nc = *q;
                if (ignitable(nc & 7)) *q = (CellT)((nc & ~7) | ST_BURNING);  // fire.py:779 (several sprites may write it: same value)
            }
    }
}
__global__ void k_cs_clock(DevParams p, int par) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= p.E) return;
    EnvMeta m = p.meta[(long long)par * p.meta_stride + env];
    if (m.running) m.t += 1;  // durations += 1 (fire.py:785); this manager has no GameStatus
    m.any_live = m.any_cand = 0;
    p.meta[(long long)(par ^ 1) * p.meta_stride + env] = m;
}

__global__ void k_mark_envs(uint8_t* mark, const int32_t* envs, int env0, int n) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) mark[envs ? envs[k] : env0 + k] = 1;
}

// ---------------------------------------------------------------------------------------
// dispatch on the cell width
// ---------------------------------------------------------------------------------------
#define DISPATCH(s, KERNEL, grid, block, ...)                                              \
    do {                                                                                   \
        if ((s)->cell_bytes == 1) SFB_LAUNCH(KERNEL<uint8_t>, (grid), (block), 0, (s)->stream, __VA_ARGS__); \
        else SFB_LAUNCH(KERNEL<uint16_t>, (grid), (block), 0, (s)->stream, __VA_ARGS__);          \
        (s)->launches_all++;                                                               \
    } while (0)

static unsigned cap_grid(sfb_sim* s, long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    long long cap = (long long)s->n_sm * 32;
    return (unsigned)std::max(1LL, std::min(b, cap));
}

// ---------------------------------------------------------------------------------------
// lifetime
// ---------------------------------------------------------------------------------------
extern "C" const char* sfb_last_error(void) { return g_err; }
extern "C" int sfb_abi_version(void) { return SFB_ABI_VERSION; }

extern "C" void sfb_destroy(sfb_sim* s) {
    if (!s) return;
    cudaSetDevice(s->prm.device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    s->stream = s->own_stream;
    cudaFree(s->d.state);
    cudaFree(s->d.burn);
    if (s->d.ros_w != s->d.ros) cudaFree(s->d.ros_w);
    cudaFree(s->d.ros);
    cudaFree(s->d.ign);
    cudaFree((void*)s->d.stat);
    cudaFree((void*)s->d.drv);
    cudaFree(s->d.meta);
    cudaFree(s->d.queue);
    cudaFree(s->d.rows);
    cudaFree(s->d.unit_act);
    cudaFree(s->d.units);
    cudaFree(s->d.bits);
    cudaFree(s->d.tile_act);
    cudaFree(s->d.wl[0]);
    cudaFree(s->d.wl[1]);
    cudaFree(s->d.listed);
    cudaFree((void*)s->d.rtab);
    cudaFree(s->d.ros_items);
    cudaFree(s->d.late);
    cudaFree(s->list_ctr);
    cudaFree(s->env_mark);
    s->groups.push_back(s->all);
    for (auto& gr : s->groups) {
        cudaFree(gr.counters);
        if (gr.stream) cudaStreamDestroy(gr.stream);
        if (gr.stream_prio) cudaStreamDestroy(gr.stream_prio);
        if (gr.done) cudaEventDestroy(gr.done);
    }
    if (s->fork_ev) cudaEventDestroy(s->fork_ev);
    if (s->pair_graph) cudaGraphExecDestroy(s->pair_graph);
    if (s->group_graph) cudaGraphExecDestroy(s->group_graph);
    for (auto& m : s->log_mapped)
        if (m) cudaFreeHost(m);
    cudaFree(s->log_counts);
    if (s->log_head) cudaFreeHost(s->log_head);
    delete s->pool;
    cudaFree((void*)s->d.filler);
    cudaFree(s->stage);
    cudaFree(s->obs);
    cudaFree(s->small);
    for (int i = 0; i < 2; ++i) {
        if (s->pts_host[i]) cudaFreeHost(s->pts_host[i]);
        if (s->pts_ev[i]) cudaEventDestroy(s->pts_ev[i]);
    }
    for (auto& e : s->ev)
        if (e) cudaEventDestroy(e);
    for (auto& e : s->span)
        if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// k_tiles is compiled once per number of source planes (max_fire_duration 1 .. BITS_MAX_DUR), cell type and
// with / without the statistics counters (kernel-timing passes only)
typedef void (*tiles_fn_t)(DevParams, int, int);
template <typename CellT, bool STATS>
static tiles_fn_t tiles_kernel_of(int max_dur) {
    switch (max_dur) {
        case 1: return k_tiles<CellT, 1, STATS>;
        case 2: return k_tiles<CellT, 2, STATS>;
        case 3: return k_tiles<CellT, 3, STATS>;
        case 4: return k_tiles<CellT, 4, STATS>;
        case 5: return k_tiles<CellT, 5, STATS>;
        case 6: return k_tiles<CellT, 6, STATS>;
        default: return k_tiles<CellT, 7, STATS>;
    }
}
static tiles_fn_t tiles_kernel(int cell_bytes, int max_dur, bool stats = false) {
    if (stats) return cell_bytes == 1 ? tiles_kernel_of<uint8_t, true>(max_dur) : tiles_kernel_of<uint16_t, true>(max_dur);
    return cell_bytes == 1 ? tiles_kernel_of<uint8_t, false>(max_dur) : tiles_kernel_of<uint16_t, false>(max_dur);
}

static int create_impl(const sfb_params* prm, sfb_sim* s) {
    s->prm = *prm;
    CU(cudaSetDevice(prm->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, prm->device));
    s->n_sm = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
    s->stream = s->own_stream;
    for (auto& e : s->ev) CU(cudaEventCreate(&e));
    for (auto& e : s->span) CU(cudaEventCreate(&e));

    DevParams& d = s->d;
    memset(&d, 0, sizeof(d));
    d.H = prm->H;
    d.W = prm->W;
    d.E = prm->E;
    d.pitch = (prm->W + 15) / 16 * 16;
    d.plane = (int64_t)d.H * d.pitch;
    d.max_dur = prm->max_fire_duration;
    d.diagonal = (prm->flags & SFB_DIAGONAL_SPREAD) != 0;
    d.attenuate = (prm->flags & SFB_ATTENUATE_LINE_ROS) != 0;
    d.shared_static = (prm->flags & SFB_SHARED_STATIC) != 0;
    d.keep_ros = (prm->flags & SFB_KEEP_ROS) != 0;
    d.has_max_time = (prm->flags & SFB_HAS_MAX_TIME) != 0;
    d.ps = prm->pixel_scale;
    d.dt = prm->update_rate;
    d.max_time = prm->max_time;
    d.part = SfbParticle{prm->h, prm->S_T, prm->S_e, prm->p_p, prm->M_f};
    s->cell_bytes = (prm->max_fire_duration <= Cell<uint8_t>::M - 1 && !(prm->flags & SFB_WIDE_CELLS)) ? 1 : 2;

    const int wr = 32 * (16 / s->cell_bytes);  // cells per warp row
    d.strips = (d.pitch + wr - 1) / wr;
    const bool tma = !(prm->flags & SFB_SWEEP_LDG) && prm->slab_total_H == 0;
    int R = prm->rows_per_chunk;
    if (R <= 0) {
        // enough warps to fill the machine a few times over, as few halo rows as possible
        R = tma ? 64 : 32;
        const long long want = (long long)s->n_sm * 64;
        while (R > 8 && (long long)d.E * d.strips * ((d.H + R - 1) / R) < want) R /= 2;
        if (tma) R -= 2;  // rows_per_chunk + 2 halo rows = whole TMA boxes
    }
    if (tma)
        R = std::max(TMA_BOX_ROWS - 2, (R + 2 + TMA_BOX_ROWS - 1) / TMA_BOX_ROWS * TMA_BOX_ROWS - 2);
    else
        R = std::max(4, (R + 3) / 4 * 4);
    d.rows_per_chunk = R;
    d.chunks = (d.H + R - 1) / R;
    d.n_units = (int64_t)d.E * d.chunks * d.strips;

    const int64_t total = (int64_t)d.E * d.plane;
    // front end.  The sweep front ends (row units from 1024 units up, the dense TMA sweep below) are the
    // default: their row tasks read the state in 512-byte runs, which HBM serves at full speed.  The
    // list-driven step (sfb_lists.cuh, SFB_FRONT_LISTS) does less work per step -- one thread per cell
    // that can change -- but every one of its reads is a scattered 32-byte sector, and B200 sustains only
    // ~40 G such reads per second once they miss L2 (profiles/r02_gather_latency.txt): it wins where the
    // planes fit L2 (single envs, small batches: SFB_FRONT=lists or the flag), not at the 2048^2 x 1024 target.
    s->front_lists = prm->slab_total_H == 0 && (prm->flags & SFB_FRONT_LISTS) != 0;
    if (const char* e = getenv("SFB_FRONT")) s->front_lists = prm->slab_total_H == 0 && strcmp(e, "lists") == 0;
    // bitboard front end: needs a short sprite life (one plane per duration) and tile indices that fit the task
    // It is the default for handles big enough to skip units at all (the same threshold as row units) when
    // the caller asked for no particular front end; SFB_FRONT=rows|dense|... or any of the sweep flags keep the others.
    {
        const bool eligible = !s->front_lists && prm->slab_total_H == 0 && prm->max_fire_duration <= BITS_MAX_DUR &&
                              d.H <= 32 * 65535 && d.W <= TW * 65535;
        const bool asked_other = (prm->flags & (SFB_UNIT_SKIP_OFF | SFB_UNIT_SKIP_ON | SFB_UNIT_CHUNKS | SFB_SWEEP_LDG | SFB_STEP_GRAPH)) != 0 ||
                                 prm->rows_per_chunk > 0 || getenv("SFB_UNIT_SKIP") || getenv("SFB_UNIT_ROWS");
        const int64_t row_units = (int64_t)d.E * d.H * ((d.W + wr - 1) / wr);
        s->front_bits = eligible && ((prm->flags & SFB_FRONT_BITS) != 0 || (!asked_other && row_units >= 1024));
        if (const char* e = getenv("SFB_FRONT")) s->front_bits = eligible && strcmp(e, "bits") == 0;
    }
    s->pdl = 1;
    if (const char* e = getenv("SFB_PDL")) s->pdl = atoi(e) != 0;
    if (d.H >= (1 << LE_BITS) || d.W >= (1 << LE_BITS) || d.E >= (1 << 22)) s->front_lists = 0;  // entry fields
    // row tasks of the sweep front ends pack y into 20 bits and the strip into 8 (make_row_task)
    if (!s->front_lists && (d.H >= (1 << 20) || d.strips > 256))
        return fail(SFB_ERR_INVALID, "sfb_create: grid of %d rows x %d strips of %d columns exceeds the row-task fields (2^20 rows, 256 strips)",
                    d.H, d.strips, wr);
    int64_t qcap = prm->queue_capacity;
    if (qcap <= 0) qcap = total <= (8 << 20) ? total : std::max<int64_t>(8 << 20, total / 8);
    d.qcap = qcap;

    int rc;
    // the slab mailbox sits behind the state plane, inside the same allocation, so that the one
    // IPC handle of the plane also maps it into the peers
    s->mailbox_off = ((size_t)total * s->cell_bytes + 255) / 256 * 256;
    if ((rc = dmalloc(s, (char**)&d.state, s->mailbox_off + sizeof(SlabMailbox)))) return rc;
    d.mailbox = reinterpret_cast<SlabMailbox*>((char*)d.state + s->mailbox_off);
    CU(cudaMemsetAsync(d.mailbox, 0, sizeof(SlabMailbox), s->stream));
    if ((rc = dmalloc(s, &d.burn, (size_t)total * 8))) return rc;
    if (d.keep_ros && (rc = dmalloc(s, &d.ros, (size_t)total * 8))) return rc;
    d.ros_w = d.ros;
    if ((prm->flags & SFB_KEEP_IGNITION) && (rc = dmalloc(s, &d.ign, (size_t)total * 4))) return rc;
    const int64_t stat_cells = d.shared_static ? d.plane : total;
    if ((rc = dmalloc(s, (StaticRec**)&d.stat, (size_t)stat_cells * sizeof(StaticRec)))) return rc;
    {
        // every input of the Rothermel evaluation is static per (cell, direction): a table of the eight
        // float64 rates per static cell (64 B) replaces the evaluation in the step -- if it fits easily
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = (size_t)stat_cells * 8 * sizeof(double);
        const size_t rest = (size_t)total * 12;  // burn + lists still to be allocated
        if (!getenv("SFB_NO_RTAB") && free_b > rest && need <= (free_b - rest) / 2 && (rc = dmalloc(s, (double**)&d.rtab, need))) return rc;
    }
    if (!d.rtab && (rc = dmalloc(s, (DerivedRec**)&d.drv, (size_t)stat_cells * sizeof(DerivedRec)))) return rc;
    // (the group views copy these pointers: make_view runs after this point)
    s->static_dirty = 1;
    if ((rc = dmalloc(s, &d.meta, (size_t)2 * d.E * sizeof(EnvMeta)))) return rc;
    d.track = (prm->flags & SFB_TRACK_CHANGES) != 0;
    {
        const size_t n = (size_t)d.pitch + 32;
        std::vector<uint8_t> fill(n * s->cell_bytes, 0);
        for (size_t i = 0; i < n; ++i) fill[i * s->cell_bytes] = (uint8_t)ST_BURNED;  // little endian
        if ((rc = dmalloc(s, (char**)&d.filler, fill.size()))) return rc;
        CU(cudaMemcpy((void*)d.filler, fill.data(), fill.size(), cudaMemcpyHostToDevice));
    }

    s->use_tma = !(prm->flags & SFB_SWEEP_LDG) && prm->slab_total_H == 0;
    typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* encode = nullptr;
    if (s->use_tma) {
        // driver entry point through the runtime: no link-time dependency on libcuda
        cudaDriverEntryPointQueryResult qres;
        CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &encode, cudaEnableDefault, &qres));
        if (!encode || qres != cudaDriverEntryPointSuccess)
            return fail(SFB_ERR_CUDA, "sfb_create: cuTensorMapEncodeTiled is not available in this driver");
        CU(cudaFuncSetAttribute(k_sweep_tma<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_BLOCK_SMEM));
        CU(cudaFuncSetAttribute(k_sweep_tma<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_BLOCK_SMEM));
    }
    int sweep_per_sm = 0, rows_per_sm = 0;
    if (s->use_tma) {
        if (s->cell_bytes == 1)
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sweep_per_sm, k_sweep_tma<uint8_t>, SWEEP_WARPS * 32, TMA_BLOCK_SMEM));
        else
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sweep_per_sm, k_sweep_tma<uint16_t>, SWEEP_WARPS * 32, TMA_BLOCK_SMEM));
    } else {
        if (s->cell_bytes == 1)
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sweep_per_sm, k_sweep_ldg<uint8_t>, SWEEP_WARPS * 32, 0));
        else
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&sweep_per_sm, k_sweep_ldg<uint16_t>, SWEEP_WARPS * 32, 0));
    }
    if (s->cell_bytes == 1) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rows_per_sm, k_rows<uint8_t>, ROWS_WARPS * 32, 0));
    else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rows_per_sm, k_rows<uint16_t>, ROWS_WARPS * 32, 0));
    if (sweep_per_sm < 1 || rows_per_sm < 1) return fail(SFB_ERR_CUDA, "sfb_create: a step kernel does not fit on an SM");

    // env groups
    int G = prm->env_groups;
    if (G <= 0) G = prm->slab_total_H != 0 ? 1 : (d.E >= 512 ? 4 : (d.E >= 64 ? 2 : 1));
    if (prm->slab_total_H != 0) G = 1;
    // bitboard handles step all envs as one group: a step is two short kernels whose tails are not worth hiding
    // behind another group's launches (measured on the target batch with the three-kernel form: 1 group 128 T,
    // 2 groups 117 T, 4 groups 67 T cell-updates/s), and one group means one tile list for the setup kernels
    if (s->front_bits) G = 1;
    G = std::min(G, std::min(d.E, 16));
    d.meta_stride = d.E;
    d.idx_base = 0;
    // every warp-row of the grid (every tile of a bitboard handle): the list cannot overflow
    const int64_t tasks_per_env = s->front_bits ? std::max<int64_t>((int64_t)d.H * d.strips, (int64_t)((d.H + 31) / 32) * ((d.W + TW - 1) / TW))
                                                : (int64_t)d.H * d.strips;
    d.rows_cap = (int64_t)d.E * tasks_per_env;
    if (s->front_lists) {
        G = 1;
        d.rows_cap = 1;
        // every listed cell has one entry: a list as long as the grid cannot overflow; big batches get an
        // eighth of that (an overflow turns the handle to the dense form of the step, results unchanged)
        d.wl_cap = prm->queue_capacity > 0 ? prm->queue_capacity : (total <= (32 << 20) ? total : std::max<int64_t>(32 << 20, total / 8));
        for (int k = 0; k < 2; ++k)
            if ((rc = dmalloc(s, &d.wl[k], (size_t)d.wl_cap * 8))) return rc;
        const size_t words = (size_t)(total + 31) / 32 + 1;
        if ((rc = dmalloc(s, &d.listed, words * 4))) return rc;
        CU(cudaMemsetAsync(d.listed, 0, words * 4, s->stream));
        if ((rc = dmalloc(s, &s->list_ctr, 16 * sizeof(unsigned long long)))) return rc;
        CU(cudaMemsetAsync(s->list_ctr, 0, 16 * sizeof(unsigned long long), s->stream));
        d.front_stats = s->list_ctr + 8;
        d.late_count = reinterpret_cast<unsigned int*>(s->list_ctr + 6);
        d.late_cap = 1 << 16;
        if ((rc = dmalloc(s, &d.late, (size_t)d.late_cap * 8))) return rc;
        d.wl_count = s->list_ctr;
        d.ros_count = s->list_ctr + 2;
        d.broken = reinterpret_cast<int32_t*>(s->list_ctr + 3);
        d.ticket = reinterpret_cast<unsigned int*>(s->list_ctr + 4);
        d.dense_now = reinterpret_cast<int32_t*>(s->list_ctr + 5);
        if ((rc = dmalloc(s, &s->env_mark, (size_t)d.E))) return rc;
        CU(cudaMemsetAsync(s->env_mark, 0, (size_t)d.E, s->stream));
        if (d.keep_ros) {
            d.ros_cap = total;
            if ((rc = dmalloc(s, &d.ros_items, (size_t)d.ros_cap * 16))) return rc;
        }
#define SFB_FRONT_SETUP(K)                                                                       \
    do {                                                                                        \
        CU(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, FRONT_SMEM));   \
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&s->front_blocks, K, FRONT_THREADS, FRONT_SMEM)); \
    } while (0)
        if (s->cell_bytes == 1 && d.rtab) SFB_FRONT_SETUP((k_front<uint8_t, true>));
        else if (s->cell_bytes == 1) SFB_FRONT_SETUP((k_front<uint8_t, false>));
        else if (d.rtab) SFB_FRONT_SETUP((k_front<uint16_t, true>));
        else SFB_FRONT_SETUP((k_front<uint16_t, false>));
#undef SFB_FRONT_SETUP
        if (s->front_blocks < 1) return fail(SFB_ERR_CUDA, "sfb_create: k_front does not fit on an SM");
        s->front_blocks *= s->n_sm;
        if (const char* e = getenv("SFB_FRONT_BLOCKS")) s->front_blocks = std::max(1, atoi(e));
    } else {
        if ((rc = dmalloc(s, &d.queue, (size_t)d.qcap * 8))) return rc;
        // (bitboard handles: two tile lists, one per step parity)
        if ((rc = dmalloc(s, &d.rows, (size_t)d.rows_cap * 8 * (s->front_bits ? 2 : 1)))) return rc;
    }
    // unit skipping: on for handles with enough units to make a list worth its launch, never in slab
    // mode (a neighbour slab's fire enters through the halo rows, which nobody here would flag)
    s->unit_skip = !s->front_lists && !s->front_bits && prm->slab_total_H == 0 && d.n_units < ((int64_t)1 << 31) &&
                   ((prm->flags & SFB_UNIT_SKIP_ON) || (!(prm->flags & SFB_UNIT_SKIP_OFF) && d.n_units >= 1024));
    if (const char* e = getenv("SFB_UNIT_SKIP"))
        if (!s->front_lists && !s->front_bits && prm->slab_total_H == 0 && d.n_units < ((int64_t)1 << 31)) s->unit_skip = atoi(e) != 0;
    s->unit_rows = s->unit_skip && !(prm->flags & SFB_UNIT_CHUNKS);
    if (const char* e = getenv("SFB_UNIT_ROWS")) s->unit_rows = s->unit_skip && atoi(e) != 0;
    d.unit_stride = (int64_t)d.chunks * d.strips;
    if (s->unit_rows) {
        // one row of one strip per unit: the flagged units are the row tasks, the state is never swept
        d.unit_rows = 1;
        d.rows_per_chunk = 1;
        d.chunks = d.H;
        d.n_units = (int64_t)d.E * d.H * d.strips;
        d.unit_stride = ((int64_t)d.H * d.strips + 3) / 4 * 4;
    }
    if (s->unit_skip) {
        const size_t flag_bytes = (size_t)d.E * d.unit_stride;
        if ((rc = dmalloc(s, &d.unit_act, flag_bytes))) return rc;
        if (!s->unit_rows && (rc = dmalloc(s, &d.units, (size_t)d.n_units * sizeof(uint32_t)))) return rc;
        CU(cudaMemsetAsync(d.unit_act, 0, flag_bytes, s->stream));
    }
    if (s->front_bits) {
        d.ring = prm->max_fire_duration + 1;
        d.tiles_x = (d.W + TW - 1) / TW;
        d.tiles_y = (d.H + 31) / 32;
        d.bits_plane = (int64_t)d.tiles_x * d.H;
        d.bits_env = (int64_t)(2 + d.ring) * d.bits_plane;
        if (d.bits_env >= ((int64_t)1 << 31)) return fail(SFB_ERR_INVALID, "sfb_create: bitboard planes of %lld words per env exceed 32-bit offsets", (long long)d.bits_env);
        d.tile_stride = ((int64_t)d.tiles_y * d.tiles_x + 15) / 16 * 16;
        d.tile_buf = (int64_t)d.E * d.tile_stride;
        if ((rc = dmalloc(s, &d.bits, (size_t)d.E * d.bits_env * 4))) return rc;
        if ((rc = dmalloc(s, &d.tile_act, (size_t)2 * d.tile_buf))) return rc;
        CU(cudaMemsetAsync(d.bits, 0, (size_t)d.E * d.bits_env * 4, s->stream));
        CU(cudaMemsetAsync(d.tile_act, 0, (size_t)2 * d.tile_buf, s->stream));
        if (d.keep_ros && (rc = dmalloc(s, &d.ros_w, (size_t)total * 8))) return rc;
        if ((rc = dmalloc(s, &s->list_ctr, 16 * sizeof(unsigned long long)))) return rc;  // ticket, statistics of timed passes
        CU(cudaMemsetAsync(s->list_ctr, 0, 16 * sizeof(unsigned long long), s->stream));
        d.ticket = reinterpret_cast<unsigned int*>(s->list_ctr + 4);
        // one kernel per step (k_tiles' last block closes the step) while no control line can exist: built, parity-
        // tested and measured SLOWER than k_tiles + k_eval on the target batch (138-141 vs 150-151 T cell-updates/s:
        // every thread's __threadfence before the ticket and one block advancing 1024 clocks cost more than the 4 us
        // of a second launch) -> opt-in
        s->fuse_eval = 0;
        if (const char* e = getenv("SFB_FUSE_EVAL")) s->fuse_eval = atoi(e) != 0;
        // (the row-task list allocated above is sized for the tile tasks)
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tiles_kernel(s->cell_bytes, prm->max_fire_duration), TILES_WARPS * 32, 0));
        if (per_sm < 1) return fail(SFB_ERR_CUDA, "sfb_create: k_tiles does not fit on an SM");
        s->tiles_fn = tiles_kernel(s->cell_bytes, prm->max_fire_duration);
        s->tiles_fn_stats = tiles_kernel(s->cell_bytes, prm->max_fire_duration, true);
        s->tiles_blocks = per_sm * s->n_sm;
    }
    CU(cudaEventCreateWithFlags(&s->fork_ev, cudaEventDisableTiming));
    // measured (profiles/r01b_bench_target_groupgraph.json): the graph joins the groups after every pair of
    // steps, which costs more overlap between groups than the saved launches give back -> opt-in
    s->patch_parallel_min = 16384;
    if (const char* e = getenv("SFB_PATCH_PARALLEL_MIN")) s->patch_parallel_min = std::max(1, atoi(e));  // tests
    s->group_graph_on = (prm->flags & SFB_STEP_GRAPH) ? 1 : 0;
    if (const char* e = getenv("SFB_GROUP_GRAPH")) s->group_graph_on = atoi(e) != 0;
    s->group_graph_steps = 2;
    if (const char* e = getenv("SFB_GROUP_GRAPH_STEPS")) s->group_graph_steps = std::max(2, atoi(e) / 2 * 2);

    // second set of streams with descending priority: the kernels of earlier groups are scheduled
    // first, so the groups finish one after the other and the host can patch the change log of a
    // finished group while the later ones still compute (used when the change log is on; without
    // it the equal-priority streams overlap the groups slightly better)
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));  // hi is numerically smaller
    int next_prio = prio_hi;
    auto make_view = [&](EnvGroup& gr, int e0, int cnt, int64_t q_off, int64_t q_cap, bool own_stream) -> int {
        const int64_t off = (int64_t)e0 * d.plane;
        gr.d = d;
        DevParams& v = gr.d;
        v.E = cnt;
        v.state = (char*)d.state + off * s->cell_bytes;
        v.burn = d.burn + off;
        if (d.ros) v.ros = d.ros + off;
        if (d.ros_w) v.ros_w = d.ros_w + off;
        if (d.ign) v.ign = d.ign + off;
        if (!d.shared_static) {
            v.stat = d.stat + off;
            if (d.drv) v.drv = d.drv + off;
            if (d.rtab) v.rtab = d.rtab + off * 8;
        }
        v.meta = d.meta + e0;
        v.idx_base = off;
        v.n_units = (int64_t)cnt * d.chunks * d.strips;
        v.queue = d.queue + q_off;
        v.qcap = q_cap;
        v.rows = d.rows + (int64_t)e0 * tasks_per_env;
        v.rows_cap = (int64_t)cnt * tasks_per_env;
        if (d.bits) {
            v.bits = d.bits + (int64_t)e0 * d.bits_env;
            v.tile_act = d.tile_act + (int64_t)e0 * d.tile_stride;
        }
        if (d.unit_act) {
            v.unit_act = d.unit_act + (int64_t)e0 * d.unit_stride;
            if (d.units) v.units = d.units + (int64_t)e0 * d.chunks * d.strips;
        }
        int rc2;
        if ((rc2 = dmalloc(s, &gr.counters, N_COUNTERS * sizeof(unsigned long long)))) return rc2;
        CU(cudaMemsetAsync(gr.counters, 0, N_COUNTERS * sizeof(unsigned long long), s->stream));
        v.qcount = gr.counters;
        v.unit_next = gr.counters + 2;
        v.rows_count = gr.counters + 4;
        v.rows_next = gr.counters + 6;
        v.overflow = reinterpret_cast<int32_t*>(gr.counters + 8);
        v.units_count = (d.unit_act || d.bits) ? gr.counters + 10 : nullptr;
        if (s->use_tma) {
            const cuuint64_t row_bytes = (cuuint64_t)d.pitch * s->cell_bytes;
            const cuuint64_t dims[3] = {row_bytes / 4, (cuuint64_t)d.H, (cuuint64_t)cnt};
            const cuuint64_t strides[2] = {row_bytes, row_bytes * (cuuint64_t)d.H};  // bytes, dims 1 and 2
            const cuuint32_t box[3] = {TMA_ROW_BYTES / 4, TMA_BOX_ROWS, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = ((encode_fn)encode)(&gr.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, v.state, dims, strides, box, estr,
                                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) return fail(SFB_ERR_CUDA, "sfb_create: cuTensorMapEncodeTiled failed (%d)", (int)r);
        }
        gr.stream = nullptr;
        gr.stream_prio = nullptr;
        gr.last_stream = nullptr;
        if (own_stream) {
            CU(cudaStreamCreateWithFlags(&gr.stream, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithPriority(&gr.stream_prio, cudaStreamNonBlocking, next_prio));
            if (next_prio < prio_lo) ++next_prio;
        }
        CU(cudaEventCreateWithFlags(&gr.done, cudaEventDisableTiming));
        const long long need = (v.n_units + SWEEP_WARPS - 1) / SWEEP_WARPS;
        // The TMA sweep would fit four blocks per SM, but their rings would take all of the SM's
        // shared memory and keep the k_rows / k_eval blocks of the other env groups out.  Two blocks
        // (8 warps x 3 boxes = 104 KB in flight per SM) already stream at full HBM speed and leave
        // room for them to co-reside: 0.764 -> 0.652 ms per step at the target.
        int per_sm = (s->use_tma && G > 1) ? std::min(sweep_per_sm, 2) : sweep_per_sm;  // one group: nothing to co-reside with
        if (const char* e = getenv("SFB_SWEEP_BLOCKS_PER_SM")) per_sm = std::max(1, std::min(sweep_per_sm, atoi(e)));
        gr.sweep_blocks = (int)std::min<long long>(need, (long long)per_sm * s->n_sm);
        const long long rows_need = (v.rows_cap + ROWS_WARPS - 1) / ROWS_WARPS;
        gr.rows_blocks = (int)std::max<long long>(1, std::min<long long>(rows_need, (long long)rows_per_sm * s->n_sm));
        const long long unit_items = d.unit_rows ? (long long)cnt * d.unit_stride / 4 : v.n_units;  // words / flags
        gr.units_blocks = (int)std::max<long long>(1, std::min<long long>((unit_items + 255) / 256, (long long)(d.unit_rows ? 8 : 4) * s->n_sm));
        return 0;
    };
    if ((rc = make_view(s->all, 0, d.E, 0, d.qcap, false))) return rc;
    s->groups.resize(G > 1 ? G : 0);
    {
        int64_t q_off = 0;
        for (int g = 0; g < (int)s->groups.size(); ++g) {
            const int e0 = (int)((long long)d.E * g / G), cnt = (int)((long long)d.E * (g + 1) / G) - e0;
            const int64_t q_cap = g + 1 == G ? d.qcap - q_off : std::max<int64_t>(1, d.qcap * cnt / d.E);
            if ((rc = make_view(s->groups[g], e0, cnt, q_off, q_cap, true))) return rc;
            q_off += q_cap;
        }
    }
    // change logs: one per env group (one in all if there is a single group).  They live in pinned
    // HOST memory mapped into the device address space: the warp-aggregated appends of k_eval
    // are coalesced 256-byte posted writes over PCIe, so by the time a group's stream is idle its
    // entries are already on the host and no D2H copy is needed.
    if (d.track) {
        const int nl = std::max<int>(1, (int)s->groups.size());
        const int64_t cap_all = std::min<int64_t>(std::max<int64_t>(1 << 20, total / 16), (int64_t)16 << 20);
        if ((rc = dmalloc(s, &s->log_counts, (size_t)nl * 2 * sizeof(unsigned long long)))) return rc;
        CU(cudaMemsetAsync(s->log_counts, 0, (size_t)nl * 2 * sizeof(unsigned long long), s->stream));
        CU(cudaMallocHost((void**)&s->log_head, (size_t)nl * 2 * sizeof(unsigned long long)));
        d.n_logs = nl;
        for (int g = 0; g < nl; ++g) {
            const int e0 = nl == 1 ? 0 : (int)((long long)d.E * g / nl);
            const int e1 = nl == 1 ? d.E : (int)((long long)d.E * (g + 1) / nl);
            LogRef& L = d.logs[g];
            L.cap = std::max<int64_t>(1 << 16, cap_all * (e1 - e0) / d.E);
            CU(cudaHostAlloc((void**)&s->log_mapped[g], (size_t)L.cap * 8, cudaHostAllocMapped | cudaHostAllocPortable));
            CU(cudaHostGetDevicePointer((void**)&L.buf, s->log_mapped[g], 0));
            L.count = s->log_counts + 2 * g;
            d.log_e0[g] = e0;
            d.log_e0[g + 1] = e1;
        }
        auto bind = [&](EnvGroup& gr, int g) {
            gr.d.n_logs = d.n_logs;
            for (int k = 0; k <= nl; ++k) gr.d.log_e0[k] = d.log_e0[k];
            for (int k = 0; k < nl; ++k) gr.d.logs[k] = d.logs[k];
            gr.d.chg = d.logs[g].buf;
            gr.d.chg_count = d.logs[g].count;
            gr.d.chg_overflow = reinterpret_cast<int32_t*>(d.logs[g].count + 1);
            gr.d.chg_cap = d.logs[g].cap;
        };
        bind(s->all, 0);  // only used for stepping when there is a single log
        for (int g = 0; g < (int)s->groups.size(); ++g) bind(s->groups[g], g);
        d.chg = d.logs[0].buf;
        d.chg_count = d.logs[0].count;
        d.chg_overflow = reinterpret_cast<int32_t*>(d.logs[0].count + 1);
        d.chg_cap = d.logs[0].cap;
    }
    // the handle-wide params point at the `all` view's counters (setup kernels never touch them)
    d.qcount = s->all.d.qcount;
    d.unit_next = s->all.d.unit_next;
    d.rows_count = s->all.d.rows_count;
    d.rows_next = s->all.d.rows_next;
    d.overflow = s->all.d.overflow;
    d.units_count = s->all.d.units_count;

    CU(cudaMemsetAsync((void*)d.stat, 0, (size_t)stat_cells * sizeof(StaticRec), s->stream));
    CU(cudaMemsetAsync(d.meta, 0, (size_t)2 * d.E * sizeof(EnvMeta), s->stream));  // running = 0
    DISPATCH(s, k_clear_envs, cap_grid(s, total, 256), 256, d, (const int32_t*)nullptr, d.E, 1);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_create(const sfb_params* prm, sfb_sim** out) {
    if (!prm || !out) return fail(SFB_ERR_INVALID, "sfb_create: null argument");
    *out = nullptr;
    if (prm->abi_version != SFB_ABI_VERSION)
        return fail(SFB_ERR_INVALID, "sfb_create: abi_version %d, library is %d", prm->abi_version, SFB_ABI_VERSION);
    if (prm->H < 1 || prm->W < 1 || prm->E < 1) return fail(SFB_ERR_INVALID, "sfb_create: H, W, E must be >= 1");
    if (prm->max_fire_duration < 1 || prm->max_fire_duration > Cell<uint16_t>::M - 1)
        return fail(SFB_ERR_INVALID, "sfb_create: max_fire_duration %d outside 1..%d", prm->max_fire_duration,
                    Cell<uint16_t>::M - 1);
    if (!(prm->update_rate == prm->update_rate) || !(prm->pixel_scale == prm->pixel_scale))
        return fail(SFB_ERR_INVALID, "sfb_create: NaN pixel_scale / update_rate");
    if ((int64_t)prm->E * prm->H * ((prm->W + 15) / 16 * 16) >= ((int64_t)1 << 47))
        return fail(SFB_ERR_INVALID, "sfb_create: more than 2^47 cells");
    if (prm->slab_total_H != 0 &&
        (prm->slab_y0 < 0 || prm->slab_y0 + prm->H > prm->slab_total_H))
        return fail(SFB_ERR_INVALID, "sfb_create: slab rows [%d, %d) outside grid of %d rows", prm->slab_y0,
                    prm->slab_y0 + prm->H, prm->slab_total_H);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SFB_ERR_CUDA, "sfb_create: no CUDA device (%s); this library has no CPU path",
                    cudaGetErrorString(e));
    if (prm->device < 0 || prm->device >= ndev) return fail(SFB_ERR_INVALID, "sfb_create: device %d of %d", prm->device, ndev);
    sfb_sim* s = new (std::nothrow) sfb_sim();
    if (!s) return fail(SFB_ERR_NOMEM, "sfb_create: host allocation failed");
    int rc = create_impl(prm, s);
    if (rc) {
        char keep[sizeof(g_err)];
        memcpy(keep, g_err, sizeof(keep));
        sfb_destroy(s);
        memcpy(g_err, keep, sizeof(keep));
        return rc;
    }
    *out = s;
    return 0;
}

// ---------------------------------------------------------------------------------------
// static inputs
// ---------------------------------------------------------------------------------------
static int set_static_range(sfb_sim* s, int env, int plane, const float* dev_src) {
    const DevParams& d = s->d;
    int first = env, n = 1;
    if (d.shared_static) {
        first = 0;
        n = 1;
    } else if (env < 0) {
        first = 0;
        n = d.E;
    }
    const long long total = (long long)n * d.H * d.W;
    SFB_LAUNCH(k_set_static, cap_grid(s, total, 256), 256, 0, s->stream, d, first, n, plane, dev_src);
    s->launches_all++;
    s->static_dirty = 1;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sfb_set_static(sfb_sim* s, int32_t env, int32_t plane, const float* host) {
    if (!s || !host) return fail(SFB_ERR_INVALID, "sfb_set_static: null argument");
    if (plane < 0 || plane >= SFB_N_STATIC) return fail(SFB_ERR_INVALID, "sfb_set_static: plane %d", plane);
    if (env < -1 || env >= s->d.E) return fail(SFB_ERR_INVALID, "sfb_set_static: env %d of %d", env, s->d.E);
    int rc;
    if ((rc = use(s))) return rc;
    const size_t bytes = (size_t)s->d.H * s->d.W * sizeof(float);
    if ((rc = ensure_stage(s, bytes))) return rc;
    CU(cudaMemcpyAsync(s->stage, host, bytes, cudaMemcpyHostToDevice, s->stream));
    if ((rc = set_static_range(s, env, plane, (const float*)s->stage))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_set_static_all(sfb_sim* s, int32_t env, const float* host) {
    if (!s || !host) return fail(SFB_ERR_INVALID, "sfb_set_static_all: null argument");
    if (env < -1 || env >= s->d.E) return fail(SFB_ERR_INVALID, "sfb_set_static_all: env %d of %d", env, s->d.E);
    int rc;
    if ((rc = use(s))) return rc;
    const size_t hw = (size_t)s->d.H * s->d.W;
    if ((rc = ensure_stage(s, hw * sizeof(float) * SFB_N_STATIC))) return rc;
    CU(cudaMemcpyAsync(s->stage, host, hw * sizeof(float) * SFB_N_STATIC, cudaMemcpyHostToDevice, s->stream));
    for (int pl = 0; pl < SFB_N_STATIC; ++pl)
        if ((rc = set_static_range(s, env, pl, (const float*)s->stage + pl * hw))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_set_elevation(sfb_sim* s, int32_t env, const double* elevations) {
    if (!s || !elevations) return fail(SFB_ERR_INVALID, "sfb_set_elevation: null argument");
    if (env < -1 || env >= s->d.E) return fail(SFB_ERR_INVALID, "sfb_set_elevation: env %d of %d", env, s->d.E);
    if (s->prm.slab_total_H) return fail(SFB_ERR_STATE, "sfb_set_elevation: not available in slab mode");
    if (!(s->prm.pixel_scale != 0.0)) return fail(SFB_ERR_INVALID, "sfb_set_elevation: pixel_scale is 0");
    int rc;
    if ((rc = use(s))) return rc;
    const DevParams& d = s->d;
    const size_t bytes = (size_t)d.H * d.W * sizeof(double);
    if ((rc = ensure_stage(s, bytes))) return rc;
    CU(cudaMemcpyAsync(s->stage, elevations, bytes, cudaMemcpyHostToDevice, s->stream));
    int first = env, n = 1;
    if (d.shared_static) { first = 0; n = 1; }
    else if (env < 0) { first = 0; n = d.E; }
    SFB_LAUNCH(k_slopes, cap_grid(s, (long long)n * d.H * d.W, 256), 256, 0, s->stream, d, first, n, (const double*)s->stage, s->prm.pixel_scale);
    s->launches_all++;
    s->static_dirty = 1;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// watch-list maintenance between steps (list handles)
// ---------------------------------------------------------------------------------------
// Marks envs (a device list, or [env0, env0 + n)) and removes their entries and `listed` bits; the
// caller changes the cells, re-lists what belongs on the list and clears the marks again.
static int list_drop_envs(sfb_sim* s, const int32_t* d_envs, int env0, int n) {
    const DevParams& d = s->d;
    SFB_LAUNCH(k_mark_envs, nblocks(n, 128), 128, 0, s->stream, s->env_mark, d_envs, env0, n);
    SFB_LAUNCH(k_list_purge, cap_grid(s, std::max<int64_t>(1024, s->list_entries_last * 2 + 65536), 256), 256, 0, s->stream, d, s->lpar,
               (const uint8_t*)s->env_mark);
    CU(cudaMemsetAsync(d.wl_count + s->lpar, 0, sizeof(unsigned long long), s->stream));  // its survivors are in the other buffer now
    s->lpar ^= 1;
    SFB_LAUNCH(k_list_clear_bits, cap_grid(s, (long long)n * (d.plane / 16), 256), 256, 0, s->stream, d, d_envs, env0, n);
    s->launches_all += 3;
    CU(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------
// between-step mutations
// ---------------------------------------------------------------------------------------
extern "C" int sfb_reset(sfb_sim* s, const int32_t* envs, int32_t n, const int32_t* xy) {
    if (!s || !xy) return fail(SFB_ERR_INVALID, "sfb_reset: null argument");
    const DevParams& d = s->d;
    if (n < 1 || n > d.E) return fail(SFB_ERR_INVALID, "sfb_reset: n = %d of %d envs", n, d.E);
    const int total_H = s->prm.slab_total_H ? s->prm.slab_total_H : d.H;
    for (int k = 0; k < n; ++k) {
        if (envs && (envs[k] < 0 || envs[k] >= d.E)) return fail(SFB_ERR_INVALID, "sfb_reset: env %d of %d", envs[k], d.E);
        const int x = xy[2 * k], y = xy[2 * k + 1];
        if (x < 0 || x >= d.W || y < 0 || y >= total_H)
            return fail(SFB_ERR_INVALID, "sfb_reset: initial fire (%d, %d) outside %d x %d grid", x, y, d.W, total_H);
    }
    int rc;
    if ((rc = use(s))) return rc;
    if (s->steps_since_sync > 0) s->setup_after_step = 1;  // its log entries follow a step's
    const size_t need = (size_t)n * 3 * sizeof(int32_t);
    if ((rc = ensure_small(s, need))) return rc;
    int32_t* d_xy = s->small;
    int32_t* d_envs = envs ? s->small + 2 * (size_t)n : nullptr;
    CU(cudaMemcpyAsync(d_xy, xy, (size_t)n * 2 * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    if (envs) CU(cudaMemcpyAsync(d_envs, envs, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    if (s->front_lists && (rc = list_drop_envs(s, d_envs, 0, n))) return rc;
    if (linear_bytes(s) && d.plane % 16 == 0) {
        SFB_LAUNCH(k_clear_envs_v16, cap_grid(s, (long long)n * (d.plane / 16), 256), 256, 0, s->stream, d, (const int32_t*)d_envs, n, d.plane / 16);
        s->launches_all++;
    } else {
        DISPATCH(s, k_clear_envs, cap_grid(s, (long long)n * d.plane, 256), 256, d, (const int32_t*)d_envs, n, 1);
    }
    DISPATCH(s, k_reset_meta, nblocks(n, 128), 128, d, s->parity, (const int32_t*)d_envs, (const int32_t*)d_xy, n,
             s->prm.slab_y0, s->lpar);
    if (s->front_lists) CU(cudaMemsetAsync(s->env_mark, 0, (size_t)d.E, s->stream));
    if (s->front_bits) {  // planes of the fresh maps (one ignitable plane, one sprite of duration 0)
        DISPATCH(s, k_bits_rebuild, cap_grid(s, (long long)n * d.bits_plane, 256), 256, d, s->parity, (const int32_t*)d_envs, 0, n);
        if (!envs && n == d.E && s->ext_writes) {  // every env is fresh: no control line is left anywhere
            s->ext_writes = 0;
            s->view_epoch++;
        }
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));  // host buffers are borrowed for the call only
    return 0;
}

static void note_ext_write(sfb_sim* s);

extern "C" int sfb_apply_points(sfb_sim* s, const int32_t* pts, int64_t n) {
    if (!s || (n > 0 && !pts)) return fail(SFB_ERR_INVALID, "sfb_apply_points: null argument");
    if (n <= 0) return 0;
    const DevParams& d = s->d;
    const int total_H = s->prm.slab_total_H ? s->prm.slab_total_H : d.H;
    bool present[6] = {false, false, false, false, false, false};
    for (int64_t i = 0; i < n; ++i) {
        const int32_t* q = pts + 4 * i;
        if (q[0] < 0 || q[0] >= d.E || q[1] < 0 || q[1] >= d.W || q[2] < 0 || q[2] >= total_H || q[3] < 0 || q[3] > 5)
            return fail(SFB_ERR_INVALID, "sfb_apply_points: point %lld = (env %d, x %d, y %d, kind %d) out of range",
                        (long long)i, q[0], q[1], q[2], q[3]);
        present[q[3]] = true;
    }
    int rc;
    if ((rc = use(s))) return rc;
    if (s->steps_since_sync > 0) s->setup_after_step = 1;  // its log entries follow a step's
    note_ext_write(s);
    const size_t bytes = (size_t)n * 4 * sizeof(int32_t);
    if ((rc = ensure_small(s, bytes))) return rc;
    const bool staged = bytes <= PTS_STAGE_BYTES;  // per-step mitigation points: no device round trip
    if (staged) {
        const int slot = s->pts_slot ^= 1;
        if (!s->pts_host[slot]) {
            CU(cudaMallocHost((void**)&s->pts_host[slot], PTS_STAGE_BYTES));
            CU(cudaEventCreateWithFlags(&s->pts_ev[slot], cudaEventDisableTiming));
        }
        CU(cudaEventSynchronize(s->pts_ev[slot]));  // the copy that used this slot two calls ago
        memcpy(s->pts_host[slot], pts, bytes);
        pts = s->pts_host[slot];
    }
    CU(cudaMemcpyAsync(s->small, pts, bytes, cudaMemcpyHostToDevice, s->stream));
    // one launch per kind in BurnStatus order: FireSimulation.update_mitigation applies the
    // fireline, scratchline and wetline managers in that order (simulation.py:468-478), so
    // when two points name one cell the later kind wins, as it does there
    for (int kind = 0; kind <= 5; ++kind) {
        if (!present[kind]) continue;
        DISPATCH(s, k_apply_points, nblocks(n, 256), 256, d, (const int32_t*)s->small, (long long)n, kind,
                 s->prm.slab_y0, s->lpar);
    }
    CU(cudaGetLastError());
    if (staged) CU(cudaEventRecord(s->pts_ev[s->pts_slot], s->stream));
    else CU(cudaStreamSynchronize(s->stream));  // `pts` is borrowed for the call only
    return 0;
}

// a status written from outside: control lines (and burning cells made ignitable again) may exist from now on
static void note_ext_write(sfb_sim* s) {
    if (s->front_bits && !s->ext_writes) {
        s->ext_writes = 1;
        s->view_epoch++;  // captured step graphs launch the other set of kernels
    }
}

static int check_env_range(sfb_sim* s, const char* who, int env0, int n) {
    if (env0 < 0 || n < 1 || env0 + n > s->d.E)
        return fail(SFB_ERR_INVALID, "%s: envs [%d, %d) of %d", who, env0, env0 + n, s->d.E);
    return 0;
}

static int upload_maps(sfb_sim* s, int env0, int n, const int8_t* maps) {
    const DevParams& d = s->d;
    s->full_resync = 1;  // the change log does not describe wholesale map replacement
    note_ext_write(s);
    const size_t bytes = (size_t)n * d.H * d.W;
    int rc;
    if ((rc = ensure_stage(s, bytes))) return rc;
    CU(cudaMemcpyAsync(s->stage, maps, bytes, cudaMemcpyHostToDevice, s->stream));
    if (linear_bytes(s)) {
        SFB_LAUNCH(k_set_map_v16, cap_grid(s, (long long)bytes / 16, 256), 256, 0, s->stream,  reinterpret_cast<uint4*>((uint8_t*)d.state + (size_t)env0 * d.plane), (const uint4*)s->stage, (long long)bytes / 16);
        s->launches_all++;
    } else {
        DISPATCH(s, k_set_map, cap_grid(s, (long long)bytes, 256), 256, d, env0, n, (const int8_t*)s->stage);
    }
    // wholesale replacement: any cell may have become a control line or ignitable next to a sprite
    if (d.unit_act)
        CU(cudaMemsetAsync(d.unit_act + (size_t)env0 * d.unit_stride, 1, (size_t)n * d.unit_stride, s->stream));
    if (s->front_bits) DISPATCH(s, k_bits_rebuild, cap_grid(s, (long long)n * d.bits_plane, 256), 256, d, s->parity, (const int32_t*)nullptr, env0, n);
    if (s->front_lists) {  // the same for the watch list: drop the envs' entries, re-list from the new state
        if ((rc = list_drop_envs(s, nullptr, env0, n))) return rc;
        if (s->cell_bytes == 1) SFB_LAUNCH(k_list_rebuild<uint8_t>, cap_grid(s, (long long)n * d.plane, 256), 256, 0, s->stream, d, s->lpar, env0, n);
        else SFB_LAUNCH(k_list_rebuild<uint16_t>, cap_grid(s, (long long)n * d.plane, 256), 256, 0, s->stream, d, s->lpar, env0, n);
        s->launches_all++;
        CU(cudaMemsetAsync(s->env_mark, 0, (size_t)d.E, s->stream));
    }
    CU(cudaGetLastError());
    return 0;
}

static int download_maps(sfb_sim* s, int env0, int n, int8_t* out) {
    const DevParams& d = s->d;
    const size_t bytes = (size_t)n * d.H * d.W;
    int rc;
    if ((rc = ensure_stage(s, bytes))) return rc;
    if (linear_bytes(s)) {
        SFB_LAUNCH(k_get_map_v16, cap_grid(s, (long long)bytes / 16, 256), 256, 0, s->stream,  reinterpret_cast<const uint4*>((const uint8_t*)d.state + (size_t)env0 * d.plane), (uint4*)s->stage, (long long)bytes / 16);
        s->launches_all++;
    } else {
        DISPATCH(s, k_get_map, cap_grid(s, (long long)bytes, 256), 256, d, env0, n, (int8_t*)s->stage);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, s->stage, bytes, cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

extern "C" int sfb_set_fire_map(sfb_sim* s, int32_t env0, int32_t n, const int8_t* maps) {
    if (!s || !maps) return fail(SFB_ERR_INVALID, "sfb_set_fire_map: null argument");
    int rc;
    if ((rc = check_env_range(s, "sfb_set_fire_map", env0, n))) return rc;
    if ((rc = use(s))) return rc;
    if ((rc = upload_maps(s, env0, n, maps))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------
// the hot path
// ---------------------------------------------------------------------------------------
static int derive_if_dirty(sfb_sim* s) {
    if (!s->static_dirty) return 0;
    const DevParams& d = s->d;
    const long long cells = d.shared_static ? d.plane : (long long)d.E * d.plane;
    SFB_LAUNCH(k_derive_static, cap_grid(s, cells, 128), 128, 0, s->stream, d, cells);
    s->launches_all++;
    s->static_dirty = 0;
    CU(cudaGetLastError());
    return 0;
}

// bitboard handles with SFB_FUSE_EVAL=1: one kernel per step while no status has been written from outside (no
// control lines: nothing for k_eval to decide) and the rate-of-spread plane is not kept
static bool step_is_fused(const sfb_sim* s, const EnvGroup& gr) {
    return gr.d.bits && s->fuse_eval && !s->ext_writes && !gr.d.keep_ros;
}

static void launch_sweep(sfb_sim* s, EnvGroup& gr, cudaStream_t st, int par) {
    if (gr.d.bits) return;  // this step's tile list was written by the step before (and by the setup kernels)
    if (gr.d.unit_rows) {  // the flagged rows are this step's row tasks: nothing is swept
        SFB_LAUNCH(k_row_list, gr.units_blocks, 256, 0, st, gr.d, par);
        s->launches_all++;
        s->launches_step++;
        return;
    }
    if (gr.d.unit_act) {  // this step's list of flagged units; the sweep draws from it
        SFB_LAUNCH(k_units, gr.units_blocks, 256, 0, st, gr.d, par);
        s->launches_all++;
        s->launches_step++;
    }
    if (s->use_tma) {
        if (s->cell_bytes == 1) SFB_LAUNCH(k_sweep_tma<uint8_t>, gr.sweep_blocks, SWEEP_WARPS * 32, TMA_BLOCK_SMEM, st, gr.tmap, gr.d, par);
        else SFB_LAUNCH(k_sweep_tma<uint16_t>, gr.sweep_blocks, SWEEP_WARPS * 32, TMA_BLOCK_SMEM, st, gr.tmap, gr.d, par);
    } else {
        if (s->cell_bytes == 1) SFB_LAUNCH(k_sweep_ldg<uint8_t>, gr.sweep_blocks, SWEEP_WARPS * 32, 0, st, gr.d, par);
        else SFB_LAUNCH(k_sweep_ldg<uint16_t>, gr.sweep_blocks, SWEEP_WARPS * 32, 0, st, gr.d, par);
    }
    s->launches_all++;
    s->launches_step++;
}
static void launch_rows(sfb_sim* s, EnvGroup& gr, cudaStream_t st, int par) {
    if (gr.d.bits) {
        if (gr.d.keep_ros) {
            SFB_LAUNCH(k_ros_zero, cap_grid(s, (long long)gr.d.E * gr.d.plane, 256), 256, 0, st, gr.d);
            s->launches_all++;
        }
        const int blocks = (int)std::max<long long>(1, std::min<long long>(((long long)gr.d.E * gr.d.tiles_y * gr.d.tiles_x + TILES_WARPS - 1) / TILES_WARPS, s->tiles_blocks));
        SFB_LAUNCH_DEP(s->pdl && !gr.d.keep_ros, gr.d.tile_stats ? s->tiles_fn_stats : s->tiles_fn, blocks, TILES_WARPS * 32, 0, st, gr.d, par,
                       step_is_fused(s, gr) ? 1 : 0);
        s->launches_all++;
        s->launches_step++;
        return;
    }
    if (s->cell_bytes == 1) SFB_LAUNCH(k_rows<uint8_t>, gr.rows_blocks, ROWS_WARPS * 32, 0, st, gr.d, par);
    else SFB_LAUNCH(k_rows<uint16_t>, gr.rows_blocks, ROWS_WARPS * 32, 0, st, gr.d, par);
    s->launches_all++;
    s->launches_step++;
}
static void launch_eval(sfb_sim* s, EnvGroup& gr, cudaStream_t st, int par) {
    if (step_is_fused(s, gr)) return;  // k_tiles closed the step itself
    if (gr.d.keep_ros && !gr.d.bits) {
        SFB_LAUNCH(k_clear_ros, cap_grid(s, (long long)gr.d.E * gr.d.plane, 256), 256, 0, st, gr.d, par);
        s->launches_all++;
    }
    const bool dep = s->pdl && gr.d.bits && !gr.d.keep_ros;
    // (bitboard handles leave k_eval only the control-line cells no fire touches and the per-env clocks; a smaller
    // grid for them was measured -- 1, 2, 4 or 8 blocks per SM: the same 4 us, it is launch and drain latency)
    const int eval_blocks = s->n_sm * 8;
    if (s->cell_bytes == 1) SFB_LAUNCH_DEP(dep, k_eval<uint8_t>, eval_blocks, 256, 0, st, gr.d, par);
    else SFB_LAUNCH_DEP(dep, k_eval<uint16_t>, eval_blocks, 256, 0, st, gr.d, par);
    s->launches_all++;
    s->launches_step++;
    if (gr.d.keep_ros && gr.d.bits) {
        SFB_LAUNCH(k_ros_commit, cap_grid(s, (long long)gr.d.E * gr.d.plane, 256), 256, 0, st, gr.d, par);
        s->launches_all++;
    }
}

// refresh the fields of the group views that setters may have changed on the handle-wide params
static void sync_one_view(sfb_sim* s, EnvGroup& gr);
static void sync_group_views(sfb_sim* s) {
    sync_one_view(s, s->all);
    for (auto& gr : s->groups) sync_one_view(s, gr);
}
static void sync_one_view(sfb_sim* s, EnvGroup& gr) {
    {
        const DevParams before = gr.d;
        // the whole-handle view of a multi-group handle does not log (its steps invalidate the logs)
        gr.d.track = (&gr == &s->all && !s->groups.empty()) ? 0 : s->d.track;
        gr.d.tile_stats = (s->front_bits && s->timing) ? s->list_ctr + 8 : nullptr;
        gr.d.halo_top = s->d.halo_top;
        gr.d.halo_bottom = s->d.halo_bottom;
        gr.d.halo_top_plane = s->d.halo_top_plane;
        gr.d.halo_bottom_plane = s->d.halo_bottom_plane;
        gr.d.mailbox = s->d.mailbox;
        gr.d.slab_rank = s->d.slab_rank;
        gr.d.slab_world = s->d.slab_world;
        for (int q = 0; q < SLAB_MAX_WORLD; ++q) gr.d.peer_box[q] = s->d.peer_box[q];
        if (memcmp(&before, &gr.d, sizeof(DevParams)) != 0) s->view_epoch++;  // captured graphs are stale
    }
}

// single-group handles (slab mode, small batches): the two halves of a step on the handle's stream
static int enqueue_sweep(sfb_sim* s) {
    int rc;
    if ((rc = derive_if_dirty(s))) return rc;
    sync_group_views(s);
    EnvGroup& gr = s->all;
    if (s->timing) CU(cudaEventRecord(s->ev[0], s->stream));
    launch_sweep(s, gr, s->stream, s->parity);
    if (s->timing) CU(cudaEventRecord(s->ev[1], s->stream));
    launch_rows(s, gr, s->stream, s->parity);
    if (s->timing) CU(cudaEventRecord(s->ev[2], s->stream));
    s->in_step = 1;
    return 0;
}

static int enqueue_eval(sfb_sim* s) {
    launch_eval(s, s->all, s->stream, s->parity);
    s->parity ^= 1;
    s->in_step = 0;
    if (s->timing) {
        CU(cudaEventRecord(s->ev[3], s->stream));
        CU(cudaEventSynchronize(s->ev[3]));
        float a = 0, b = 0, c = 0;
        CU(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
        CU(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
        CU(cudaEventElapsedTime(&c, s->ev[2], s->ev[3]));
        s->sweep_ms += a;
        s->rows_ms += b;
        s->eval_ms += c;
        s->timed_steps++;
    }
    return 0;
}

// n steps of every env.  One group: everything on the handle's stream.  Several groups: each on
// its own stream, forked from / joined to the handle's stream, so that the streaming kernel of
// one group overlaps the issue-bound kernels of another.  With kernel timing enabled the groups
// run one after the other and the per-kernel times add up.
// switching between the `all` view and the group views: the views keep separate counters, and a
// view that sat out some steps may hold counts of the wrong parity
static int enter_mode(sfb_sim* s, int mode) {
    if (s->last_mode == mode) return 0;
    if (s->last_mode != 0) {
        CU(cudaMemsetAsync(s->all.counters, 0, N_COUNTERS * sizeof(unsigned long long), s->stream));
        for (auto& gr : s->groups) CU(cudaMemsetAsync(gr.counters, 0, N_COUNTERS * sizeof(unsigned long long), s->stream));
    }
    s->last_mode = mode;
    return 0;
}

// replay (capturing it first if needed) the graph of two steps of the whole-handle view; parity must be 0
static int run_pair_graph(sfb_sim* s) {
    if (!s->pair_graph || s->pair_graph_epoch != s->view_epoch) {
        if (s->pair_graph) CU(cudaGraphExecDestroy(s->pair_graph));
        s->pair_graph = nullptr;
        const int64_t la = s->launches_all, ls = s->launches_step;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        for (int par = 0; par < 2; ++par) {
            launch_sweep(s, s->all, s->stream, par);
            launch_rows(s, s->all, s->stream, par);
            launch_eval(s, s->all, s->stream, par);
        }
        CU(cudaStreamEndCapture(s->stream, &g));
        s->pair_launches_all = s->launches_all - la;
        s->pair_launches_step = s->launches_step - ls;
        s->launches_all = la;
        s->launches_step = ls;
        cudaError_t e = cudaGraphInstantiate(&s->pair_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        s->pair_graph_epoch = s->view_epoch;
    }
    CU(cudaGraphLaunch(s->pair_graph, s->stream));
    s->launches_all += s->pair_launches_all;
    s->launches_step += s->pair_launches_step;
    return 0;
}

// n steps of every group, kernel by kernel, each group on its own stream
static int enqueue_group_steps(sfb_sim* s, int n) {
    CU(cudaEventRecord(s->fork_ev, s->stream));
    for (auto& gr : s->groups) {
        gr.last_stream = s->d.track ? gr.stream_prio : gr.stream;
        CU(cudaStreamWaitEvent(gr.last_stream, s->fork_ev, 0));
    }
    int par = s->parity;
    for (int i = 0; i < n; ++i) {
        for (auto& gr : s->groups) {
            launch_sweep(s, gr, gr.last_stream, par);
            launch_rows(s, gr, gr.last_stream, par);
            launch_eval(s, gr, gr.last_stream, par);
        }
        par ^= 1;
    }
    s->parity = par;
    const bool heads = s->d.track && s->log_head && (int)s->groups.size() == s->d.n_logs;
    for (size_t g = 0; g < s->groups.size(); ++g) {
        auto& gr = s->groups[g];
        // the group's {count, overflow} rides behind its last kernel, so that sfb_sync_fire_maps
        // finds it in host memory as soon as the group's event has fired
        if (heads)
            CU(cudaMemcpyAsync(s->log_head + 2 * g, s->d.logs[g].count, 2 * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, gr.last_stream));
        CU(cudaEventRecord(gr.done, gr.last_stream));
        CU(cudaStreamWaitEvent(s->stream, gr.done, 0));
    }
    s->head_valid = heads;
    return 0;
}

// replay (capturing it first if needed) two steps of every env group as one graph; parity must be 0
static int run_group_graph(sfb_sim* s) {
    if (!s->group_graph || s->group_graph_epoch != s->view_epoch) {
        if (s->group_graph) CU(cudaGraphExecDestroy(s->group_graph));
        s->group_graph = nullptr;
        const int64_t la = s->launches_all, ls = s->launches_step;
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
        CU(cudaEventRecord(s->fork_ev, s->stream));
        for (auto& gr : s->groups) CU(cudaStreamWaitEvent(gr.stream, s->fork_ev, 0));
        for (int k = 0; k < s->group_graph_steps; ++k)
            for (auto& gr : s->groups) {
                launch_sweep(s, gr, gr.stream, k & 1);
                launch_rows(s, gr, gr.stream, k & 1);
                launch_eval(s, gr, gr.stream, k & 1);
            }
        for (auto& gr : s->groups) {
            CU(cudaEventRecord(gr.done, gr.stream));
            CU(cudaStreamWaitEvent(s->stream, gr.done, 0));
        }
        CU(cudaStreamEndCapture(s->stream, &g));
        s->group_launches_all = s->launches_all - la;
        s->group_launches_step = s->launches_step - ls;
        s->launches_all = la;
        s->launches_step = ls;
        cudaError_t e = cudaGraphInstantiate(&s->group_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) return fail(SFB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        s->group_graph_epoch = s->view_epoch;
    }
    CU(cudaGraphLaunch(s->group_graph, s->stream));
    s->launches_all += s->group_launches_all;
    s->launches_step += s->group_launches_step;
    return 0;
}

// list handles: one k_front per step on the handle's stream (k_tail behind it with attenuation; the
// rate_of_spread plane is rebuilt in between for SFB_KEEP_ROS handles)
static int enqueue_list_steps(sfb_sim* s, int n) {
    int rc;
    if ((rc = derive_if_dirty(s))) return rc;
    const DevParams& d = s->d;
    for (int i = 0; i < n; ++i) {
#if !defined(SFB_EMU) && defined(SFB_EXPERIMENT_SORT)
        if (getenv("SFB_SORT_LIST")) {  // experiment: what an (env, y, x)-ordered list would buy
            unsigned long long n_host = 0;
            CU(cudaMemcpyAsync(&n_host, d.wl_count + s->lpar, 8, cudaMemcpyDeviceToHost, s->stream));
            CU(cudaStreamSynchronize(s->stream));
            n_host = std::min<unsigned long long>(n_host, (unsigned long long)d.wl_cap);
            static void* tmp = nullptr;
            static size_t tmp_bytes = 0;
            size_t need = 0;
            cub::DeviceRadixSort::SortKeys(nullptr, need, d.wl[s->lpar], d.wl[s->lpar ^ 1], (int)n_host, 0, 62, s->stream);
            if (need > tmp_bytes) { cudaFree(tmp); CU(cudaMalloc(&tmp, need)); tmp_bytes = need; }
            cub::DeviceRadixSort::SortKeys(tmp, need, d.wl[s->lpar], d.wl[s->lpar ^ 1], (int)n_host, 0, 62, s->stream);
            CU(cudaMemcpyAsync(d.wl[s->lpar], d.wl[s->lpar ^ 1], n_host * 8, cudaMemcpyDeviceToDevice, s->stream));
        }
#endif
        if (s->timing) CU(cudaEventRecord(s->ev[0], s->stream));
        if (d.keep_ros) CU(cudaMemsetAsync(d.ros_count, 0, sizeof(unsigned long long), s->stream));
        if (s->cell_bytes == 1 && d.rtab) SFB_LAUNCH((k_front<uint8_t, true>), s->front_blocks, FRONT_THREADS, FRONT_SMEM, s->stream, d, s->parity, s->lpar);
        else if (s->cell_bytes == 1) SFB_LAUNCH((k_front<uint8_t, false>), s->front_blocks, FRONT_THREADS, FRONT_SMEM, s->stream, d, s->parity, s->lpar);
        else if (d.rtab) SFB_LAUNCH((k_front<uint16_t, true>), s->front_blocks, FRONT_THREADS, FRONT_SMEM, s->stream, d, s->parity, s->lpar);
        else SFB_LAUNCH((k_front<uint16_t, false>), s->front_blocks, FRONT_THREADS, FRONT_SMEM, s->stream, d, s->parity, s->lpar);
        s->launches_all++;
        s->launches_step++;
        if (s->timing) CU(cudaEventRecord(s->ev[1], s->stream));
        if (d.keep_ros) {
            SFB_LAUNCH(k_clear_ros, cap_grid(s, (long long)d.E * d.plane, 256), 256, 0, s->stream, d, s->parity);
            SFB_LAUNCH(k_ros_apply, cap_grid(s, std::max<int64_t>(1024, s->list_entries_last + 65536), 256), 256, 0, s->stream, d);
            s->launches_all += 2;
        }
        if (s->timing) CU(cudaEventRecord(s->ev[2], s->stream));
        if (d.attenuate) {
            const unsigned tail_blocks = (unsigned)std::min<int64_t>((int64_t)s->n_sm * 4, std::max<int64_t>(1, ((int64_t)d.E * d.plane + 255) / 256));
            if (s->cell_bytes == 1) SFB_LAUNCH(k_tail<uint8_t>, tail_blocks, 256, 0, s->stream, d, s->parity, s->lpar ^ 1);
            else SFB_LAUNCH(k_tail<uint16_t>, tail_blocks, 256, 0, s->stream, d, s->parity, s->lpar ^ 1);
            s->launches_all++;
            s->launches_step++;
        }
        s->parity ^= 1;
        s->lpar ^= 1;
        if (s->timing) {
            CU(cudaEventRecord(s->ev[3], s->stream));
            CU(cudaEventSynchronize(s->ev[3]));
            float a = 0, b = 0, c = 0;
            CU(cudaEventElapsedTime(&a, s->ev[0], s->ev[1]));
            CU(cudaEventElapsedTime(&b, s->ev[1], s->ev[2]));
            CU(cudaEventElapsedTime(&c, s->ev[2], s->ev[3]));
            s->sweep_ms += a;  // k_front
            s->rows_ms += b;   // (rate_of_spread plane, parity tests only)
            s->eval_ms += c;   // k_tail
            s->timed_steps++;
        }
    }
    return 0;
}

static int enqueue_steps(sfb_sim* s, int n) {
    if (s->in_step) return fail(SFB_ERR_STATE, "a step is half done: call sfb_step_eval first");
    if (n <= 0) return 0;
    s->steps_since_sync = s->steps_since_sync > (1 << 20) ? s->steps_since_sync : s->steps_since_sync + n;
    int rc;
    if (s->front_lists) return enqueue_list_steps(s, n);
    // one view of all envs on the handle's stream: single-group handles and per-kernel timing
    if (s->groups.empty() || s->timing) {
        if ((rc = enter_mode(s, 1))) return rc;
        if (!s->groups.empty() && s->d.track) s->full_resync = 1;  // the whole-handle view does not log per group
        int i = 0;
        if (!s->timing && n >= 4 && s->stream == s->own_stream) {
            if ((rc = derive_if_dirty(s))) return rc;
            sync_group_views(s);
            if (s->parity == 1) {  // the graph starts at parity 0
                if ((rc = enqueue_sweep(s))) return rc;
                if ((rc = enqueue_eval(s))) return rc;
                i = 1;
            }
            for (; i + 1 < n; i += 2)
                if ((rc = run_pair_graph(s))) return rc;
        }
        for (; i < n; ++i) {
            if ((rc = enqueue_sweep(s))) return rc;
            if ((rc = enqueue_eval(s))) return rc;
        }
        return 0;
    }
    if ((rc = enter_mode(s, 2))) return rc;
    if ((rc = derive_if_dirty(s))) return rc;
    sync_group_views(s);
    // device-resident stepping (no change log to drain group by group): pairs of steps as one graph
    if (s->group_graph_on && !s->d.track && n >= 2 * s->group_graph_steps && s->stream == s->own_stream) {
        if (s->parity == 1) {  // the graph starts at parity 0
            if ((rc = enqueue_group_steps(s, 1))) return rc;
            --n;
        }
        for (auto& gr : s->groups) gr.last_stream = gr.stream;
        for (; n >= s->group_graph_steps; n -= s->group_graph_steps)
            if ((rc = run_group_graph(s))) return rc;  // parity is 0 again after each replay
        s->head_valid = 0;
        if (n == 0) return 0;  // the graph joined every group back into the handle's stream
    }
    return enqueue_group_steps(s, n);
}

static int enqueue_step(sfb_sim* s) { return enqueue_steps(s, 1); }

extern "C" int sfb_step_sweep(sfb_sim* s) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_step_sweep: null handle");
    if (s->in_step) return fail(SFB_ERR_STATE, "sfb_step_sweep: the previous sweep has not been evaluated");
    if (!s->groups.empty()) return fail(SFB_ERR_STATE, "sfb_step_sweep: create the handle with env_groups = 1");
    if (s->front_lists) return fail(SFB_ERR_STATE, "sfb_step_sweep: not a sweep handle (create it in slab mode or with SFB_SWEEP_LDG)");
    { int rcm = enter_mode(s, 1); if (rcm) return rcm; }
    int rc;
    if ((rc = use(s))) return rc;
    s->steps_since_sync += 2;  // half-steps driven from outside: always the ordered patch path
    if ((rc = enqueue_sweep(s))) return rc;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sfb_step_eval(sfb_sim* s) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_step_eval: null handle");
    if (!s->in_step) return fail(SFB_ERR_STATE, "sfb_step_eval: no sweep in flight");
    int rc;
    if ((rc = use(s))) return rc;
    if ((rc = enqueue_eval(s))) return rc;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sfb_slab_mailbox(sfb_sim* s, void** mailbox, int64_t* offset_from_state) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_slab_mailbox: null handle");
    if (mailbox) *mailbox = (void*)s->d.mailbox;
    if (offset_from_state) *offset_from_state = (int64_t)s->mailbox_off;
    return 0;
}

extern "C" int sfb_slab_connect(sfb_sim* s, int32_t rank, int32_t world, void* const* peer_mailboxes) {
    if (!s || !peer_mailboxes) return fail(SFB_ERR_INVALID, "sfb_slab_connect: null argument");
    if (world < 1 || world > SLAB_MAX_WORLD || rank < 0 || rank >= world)
        return fail(SFB_ERR_INVALID, "sfb_slab_connect: rank %d of %d (at most %d slabs)", rank, world, SLAB_MAX_WORLD);
    if (s->d.E > SLAB_MAX_ENVS) return fail(SFB_ERR_INVALID, "sfb_slab_connect: at most %d envs in slab mode", SLAB_MAX_ENVS);
    if (s->use_tma) return fail(SFB_ERR_STATE, "sfb_slab_connect: the handle was not created in slab mode");
    int rc;
    if ((rc = use(s))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    for (int q = 0; q < world; ++q) {
        if (!peer_mailboxes[q]) return fail(SFB_ERR_INVALID, "sfb_slab_connect: mailbox of slab %d is null", q);
        s->d.peer_box[q] = reinterpret_cast<SlabMailbox*>(peer_mailboxes[q]);
    }
    s->d.slab_rank = rank;
    s->d.slab_world = world;
    s->slab_step = 0;
    return 0;
}

extern "C" int sfb_step_slab(sfb_sim* s, int32_t n_steps) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_step_slab: null handle");
    if (s->d.slab_world < 1) return fail(SFB_ERR_STATE, "sfb_step_slab: call sfb_slab_connect first");
    if (s->in_step) return fail(SFB_ERR_STATE, "sfb_step_slab: a step is half done");
    if (!s->groups.empty()) return fail(SFB_ERR_STATE, "sfb_step_slab: slab handles have one env group");
    { int rcm = enter_mode(s, 1); if (rcm) return rcm; }
    int rc;
    if ((rc = use(s))) return rc;
    s->steps_since_sync += 2;  // (see sfb_step_sweep)
    for (int i = 0; i < n_steps; ++i) {
        const uint32_t g = ++s->slab_step;
        if ((rc = enqueue_sweep(s))) return rc;
        SFB_LAUNCH(k_slab_exchange_flags, 1, 32, 0, s->stream, s->all.d, s->parity, g);
        if ((rc = enqueue_eval(s))) return rc;
        SFB_LAUNCH(k_slab_step_done, 1, 32, 0, s->stream, s->all.d, g);
        s->launches_all += 2;
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sfb_get_parity(sfb_sim* s, int32_t* parity) {
    if (!s || !parity) return fail(SFB_ERR_INVALID, "sfb_get_parity: null argument");
    *parity = s->parity;
    return 0;
}

extern "C" int sfb_flags_device(sfb_sim* s, void** flags, int64_t* n_int32) {
    if (!s || !flags || !n_int32) return fail(SFB_ERR_INVALID, "sfb_flags_device: null argument");
    *flags = (void*)(s->d.meta + (size_t)s->parity * s->d.E);
    *n_int32 = (int64_t)s->d.E * (int64_t)(sizeof(EnvMeta) / sizeof(int32_t));
    return 0;
}

extern "C" int sfb_set_stream(sfb_sim* s, void* stream) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_set_stream: null handle");
    int rc;
    if ((rc = use(s))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->stream = stream ? (cudaStream_t)stream : s->own_stream;
    return 0;
}

extern "C" int sfb_state_device(sfb_sim* s, void** state, int64_t* plane_cells, int32_t* pitch_cells, int32_t* cell_bytes) {
    if (!s || !state) return fail(SFB_ERR_INVALID, "sfb_state_device: null argument");
    *state = s->d.state;
    if (plane_cells) *plane_cells = s->d.plane;
    if (pitch_cells) *pitch_cells = s->d.pitch;
    if (cell_bytes) *cell_bytes = s->cell_bytes;
    return 0;
}

extern "C" int sfb_static_device(sfb_sim* s, void** records, int64_t* plane_cells, int32_t* pitch_cells, int32_t* n_sets) {
    if (!s || !records) return fail(SFB_ERR_INVALID, "sfb_static_device: null argument");
    int rc;
    if ((rc = use(s))) return rc;
    CU(cudaStreamSynchronize(s->stream));  // uploads in flight have landed
    *records = (void*)s->d.stat;
    if (plane_cells) *plane_cells = s->d.plane;
    if (pitch_cells) *pitch_cells = s->d.pitch;
    if (n_sets) *n_sets = s->d.shared_static ? 1 : s->d.E;
    return 0;
}

extern "C" int sfb_ipc_export(sfb_sim* s, void* handle64) {
    if (!s || !handle64) return fail(SFB_ERR_INVALID, "sfb_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    int rc;
    if ((rc = use(s))) return rc;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, s->d.state));
    memcpy(handle64, &h, sizeof(h));
    return 0;
}

extern "C" int sfb_ipc_open(int32_t device, const void* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return fail(SFB_ERR_INVALID, "sfb_ipc_open: null argument");
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    CU(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

extern "C" int sfb_ipc_close(int32_t device, void* dev_ptr) {
    if (!dev_ptr) return 0;
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

extern "C" int sfb_set_halo(sfb_sim* s, const void* top_row, int64_t top_plane_cells, const void* bottom_row,
                            int64_t bottom_plane_cells) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_set_halo: null handle");
    if (s->use_tma && (top_row || bottom_row))
        return fail(SFB_ERR_STATE, "sfb_set_halo: the handle was not created in slab mode (slab_total_H = 0)");
    int rc;
    if ((rc = use(s))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    s->d.halo_top = top_row;
    s->d.halo_top_plane = top_plane_cells;
    s->d.halo_bottom = bottom_row;
    s->d.halo_bottom_plane = bottom_plane_cells;
    return 0;
}

extern "C" int sfb_step(sfb_sim* s, int32_t n_steps, int32_t sync) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_step: null handle");
    if (n_steps < 0) return fail(SFB_ERR_INVALID, "sfb_step: n_steps %d", n_steps);
    int rc;
    if ((rc = use(s))) return rc;
    if ((rc = enqueue_steps(s, n_steps))) return rc;
    CU(cudaGetLastError());
    if (sync) CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_step_timed(sfb_sim* s, int32_t n_steps, float* ms) {
    if (!s || !ms) return fail(SFB_ERR_INVALID, "sfb_step_timed: null argument");
    if (n_steps < 0) return fail(SFB_ERR_INVALID, "sfb_step_timed: n_steps %d", n_steps);
    int rc;
    if ((rc = use(s))) return rc;
    // the handle's own events (ev[0] / ev[3] are only used by the per-kernel timing mode, which
    // synchronises after every step, so they are free here)
    const bool was_timing = s->timing != 0;
    s->timing = 0;
    CU(cudaEventRecord(s->span[0], s->stream));
    rc = enqueue_steps(s, n_steps);
    s->timing = was_timing;
    if (rc) return rc;
    CU(cudaEventRecord(s->span[1], s->stream));
    CU(cudaEventSynchronize(s->span[1]));
    CU(cudaEventElapsedTime(ms, s->span[0], s->span[1]));
    CU(cudaGetLastError());
    return 0;
}

static int read_status(sfb_sim* s, int env0, int n, int32_t* status, double* elapsed, int32_t* steps) {
    std::vector<EnvMeta> m((size_t)n);
    CU(cudaMemcpyAsync(m.data(), s->d.meta + (size_t)s->parity * s->d.E + env0, (size_t)n * sizeof(EnvMeta),
                       cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int i = 0; i < n; ++i) {
        if (status) status[i] = m[i].running ? SFB_RUNNING : SFB_QUIT;
        if (elapsed) elapsed[i] = m[i].elapsed;
        if (steps) steps[i] = m[i].t > 0 ? m[i].t - 1 : 0;
    }
    return 0;
}

extern "C" int sfb_update(sfb_sim* s, int32_t env0, int32_t n, int8_t* maps, int32_t* status) {
    if (!s || !maps) return fail(SFB_ERR_INVALID, "sfb_update: null argument");
    int rc;
    if ((rc = check_env_range(s, "sfb_update", env0, n))) return rc;
    if ((rc = use(s))) return rc;
    if ((rc = upload_maps(s, env0, n, maps))) return rc;
    if ((rc = enqueue_step(s))) return rc;
    if ((rc = download_maps(s, env0, n, maps))) return rc;
    if (status) return read_status(s, env0, n, status, nullptr, nullptr);
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_constant_spread_update(sfb_sim* s, int32_t env0, int32_t n, int8_t* maps, int32_t rate_of_spread) {
    if (!s || !maps) return fail(SFB_ERR_INVALID, "sfb_constant_spread_update: null argument");
    if (rate_of_spread < 0) return fail(SFB_ERR_INVALID, "sfb_constant_spread_update: rate_of_spread %d", rate_of_spread);
    if (s->in_step) return fail(SFB_ERR_STATE, "sfb_constant_spread_update: a step is half done");
    if (s->front_lists || s->prm.slab_total_H) return fail(SFB_ERR_STATE, "sfb_constant_spread_update: not for list or slab handles");
    int rc;
    if ((rc = check_env_range(s, "sfb_constant_spread_update", env0, n))) return rc;
    if ((rc = use(s))) return rc;
    if ((rc = upload_maps(s, env0, n, maps))) return rc;
    const DevParams& d = s->d;
    const unsigned grid = cap_grid(s, (long long)d.E * d.plane, 256);
    DISPATCH(s, k_cs_prune, grid, 256, d, s->parity);
    DISPATCH(s, k_cs_spread, grid, 256, d, s->parity, (int)rate_of_spread);
    SFB_LAUNCH(k_cs_clock, nblocks(d.E, 128), 128, 0, s->stream, d, s->parity);
    s->launches_all++;
    s->parity ^= 1;
    if (d.unit_act) CU(cudaMemsetAsync(d.unit_act, 1, (size_t)d.E * d.unit_stride, s->stream));  // cells changed behind the flags' back
    if (s->front_bits) DISPATCH(s, k_bits_rebuild, cap_grid(s, (long long)n * d.bits_plane, 256), 256, d, s->parity, (const int32_t*)nullptr, env0, n);  // ... and the planes'
    s->full_resync = 1;
    CU(cudaGetLastError());
    if ((rc = download_maps(s, env0, n, maps))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_synchronize(sfb_sim* s) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_synchronize: null handle");
    int rc;
    if ((rc = use(s))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    CU(cudaGetLastError());
    if (s->front_lists) {
        int32_t broken = 0;
        CU(cudaMemcpy(&broken, s->d.broken, sizeof(broken), cudaMemcpyDeviceToHost));
        if (broken == 2) return fail(SFB_ERR_STATE, "sfb_synchronize: more than %lld burning cells were overdrawn and re-ignited in one step", (long long)s->d.late_cap);
    }
    if (s->d.slab_world > 0) {
        int32_t err = 0;
        CU(cudaMemcpy(&err, (const void*)&s->d.mailbox->error, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) return fail(SFB_ERR_STATE, "sfb_synchronize: a slab handshake timed out (a peer slab did not reach the same step)");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------
// results
// ---------------------------------------------------------------------------------------
extern "C" int sfb_get_fire_map(sfb_sim* s, int32_t env0, int32_t n, int8_t* out) {
    if (!s || !out) return fail(SFB_ERR_INVALID, "sfb_get_fire_map: null argument");
    int rc;
    if ((rc = check_env_range(s, "sfb_get_fire_map", env0, n))) return rc;
    if ((rc = use(s))) return rc;
    if ((rc = download_maps(s, env0, n, out))) return rc;
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

// Patch entries [0, n) of the change log into the host mirror.  Entries of one cell are in
// time order in the log and must be applied in that order.  Two passes on T pool threads:
// (1) thread k splits chunk k of the log into T buckets by owner (owner = contiguous range of
// cell indices), keeping the order; (2) owner o applies bucket (0, o), (1, o), ... in chunk
// order.  Each entry is touched twice in total, whatever T is.
// [base, base + total): the cells this log can name (its env group); the owners split that range.
// single_step: the log holds between-step entries (LOG_SETUP_BIT) followed by the entries of exactly one
// step.  The step's entries name every cell at most once, so after the (few) setup entries have been
// applied in order they are patched in parallel chunks, unsorted: one pass, no buckets, no barrier.
static void apply_log(sfb_sim* s, const unsigned long long* log, long long n, int8_t* mirror, unsigned long long base,
                      unsigned long long total, bool single_step = false) {
    const DevParams& d = s->d;
    const long long hw = (long long)d.H * d.W;
    const bool linear = d.pitch == d.W;
    auto put = [&](unsigned long long idx, int st) {
        if (linear) {
            mirror[idx] = (int8_t)st;
        } else {
            const long long env = (long long)(idx / d.plane), rem = (long long)(idx - env * d.plane);
            mirror[env * hw + (rem / d.pitch) * d.W + rem % d.pitch] = (int8_t)st;
        }
    };
    auto clear_env_part = [&](long long env, unsigned long long lo, unsigned long long hi) {
        // rows of `env` whose cells fall into the owner's range [lo, hi)
        const unsigned long long e0 = (unsigned long long)env * d.plane;
        const unsigned long long a = std::max(lo, e0), b = std::min(hi, e0 + (unsigned long long)d.plane);
        if (a >= b) return;
        if (linear) {
            memset(mirror + a, 0, (size_t)(b - a));
        } else {
            for (unsigned long long i = a; i < b; ++i)
                if ((long long)((i - e0) % d.pitch) < d.W) put(i, 0);
        }
    };
    const unsigned T = (n < s->patch_parallel_min || !s->pool) ? 1u : s->pool->size();
    if (single_step && T > 1) {
        long long n_setup = 0;
        bool resets = false;
        while (n_setup < n && (log[n_setup] & LOG_SETUP_BIT)) {
            resets |= ((int)(log[n_setup] >> 48) & 7) == LOG_ENV_RESET;
            ++n_setup;
        }
        if (!resets) {  // (whole-env clears are spread over the owners by the ordered path below)
            for (long long i = 0; i < n_setup; ++i) put(log[i] & 0xFFFFFFFFFFFFull, (int)(log[i] >> 48) & 7);
            const long long m = n - n_setup;
            s->pool->run([&](unsigned k) {
                const long long i0 = n_setup + m * k / T, i1 = n_setup + m * (k + 1) / T;
                // the mirror is far larger than any cache and the entries land all over it: every store is a
                // cache miss.  Asking for the lines a few dozen entries ahead keeps many misses in flight
                // per thread instead of one (the loop is bound by memory latency, not bandwidth).
                constexpr long long AHEAD = 24;
                if (linear) {
                    for (long long i = i0; i < i1; ++i) {
                        if (i + AHEAD < i1) __builtin_prefetch(mirror + (log[i + AHEAD] & 0xFFFFFFFFFFFFull), 1, 0);
                        mirror[log[i] & 0xFFFFFFFFFFFFull] = (int8_t)((log[i] >> 48) & 7);
                    }
                } else {
                    for (long long i = i0; i < i1; ++i) put(log[i] & 0xFFFFFFFFFFFFull, (int)(log[i] >> 48) & 7);
                }
            });
            return;
        }
    }
    if (T == 1) {
        for (long long i = 0; i < n; ++i) {
            const unsigned long long e = log[i], idx = e & 0xFFFFFFFFFFFFull;
            const int st = (int)(e >> 48) & 7;
            if (st == LOG_ENV_RESET) clear_env_part((long long)idx, base, base + total);
            else put(idx, st);
        }
        return;
    }
    const unsigned long long span = (total + T - 1) / T;  // cells per owner
    if (s->buckets.size() != (size_t)T * T) s->buckets.assign((size_t)T * T, {});
    s->pool->run([&](unsigned k) {
        // pass 1: chunk k of the log -> buckets (k, owner)
        for (unsigned o = 0; o < T; ++o) s->buckets[(size_t)k * T + o].clear();
        const long long i0 = n * k / T, i1 = n * (k + 1) / T;
        for (long long i = i0; i < i1; ++i) {
            const unsigned long long e = log[i], idx = e & 0xFFFFFFFFFFFFull;
            if (((int)(e >> 48) & 7) == LOG_ENV_RESET) {
                const unsigned long long e0 = idx * (unsigned long long)d.plane - base;  // relative to the log's range
                for (unsigned o = (unsigned)(e0 / span); o < T && (unsigned long long)o * span < e0 + (unsigned long long)d.plane; ++o)
                    s->buckets[(size_t)k * T + o].push_back(e);
            } else {
                s->buckets[(size_t)k * T + std::min<unsigned>(T - 1, (unsigned)((idx - base) / span))].push_back(e);
            }
        }
        s->pool->barrier();
        // pass 2: this thread owns cells [lo, hi)
        const unsigned o = k;
        const unsigned long long lo = base + (unsigned long long)o * span, hi = std::min(base + total, lo + span);
        for (unsigned c = 0; c < T; ++c)
            for (const unsigned long long e : s->buckets[(size_t)c * T + o]) {
                const unsigned long long idx = e & 0xFFFFFFFFFFFFull;
                const int st = (int)(e >> 48) & 7;
                if (st == LOG_ENV_RESET) clear_env_part((long long)idx, lo, hi);
                else put(idx, st);
            }
    });
}

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" int sfb_sync_fire_maps(sfb_sim* s, int8_t* mirror, int64_t* n_changes) {
    if (!s || !mirror) return fail(SFB_ERR_INVALID, "sfb_sync_fire_maps: null argument");
    int rc;
    const bool heads_ready = s->head_valid != 0;  // nothing ran since the grouped step that fetched them
    if ((rc = use(s))) return rc;
    DevParams& d = s->d;
    static const bool debug = getenv("SFB_DEBUG_TIMING") != nullptr;
    const double t0 = debug ? now_ms() : 0.0;
    const int nl = d.track ? d.n_logs : 0;
    const bool grouped = !s->groups.empty() && s->last_mode == 2;
    const bool single_step = s->steps_since_sync == 1 && !s->setup_after_step;
    s->steps_since_sync = 0;
    s->setup_after_step = 0;

    if (!d.track || s->full_resync || s->mirror != mirror) {
        CU(cudaStreamSynchronize(s->stream));
        if ((rc = download_maps(s, 0, d.E, mirror))) return rc;
        if (s->log_counts) CU(cudaMemsetAsync(s->log_counts, 0, (size_t)d.n_logs * 2 * sizeof(unsigned long long), s->stream));
        CU(cudaStreamSynchronize(s->stream));
        if (n_changes) *n_changes = -1;
        s->mirror = mirror;
        s->full_resync = 0;
        if (debug) fprintf(stderr, "[sfb_sync] full download %.3f ms\n", now_ms() - t0);
        return 0;
    }
    if (!s->pool) {
        unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        if (const char* e = getenv("SFB_HOST_THREADS")) nt = (unsigned)std::max(1, atoi(e));
        s->pool = new HostPool(std::min(nt, 64u));
    }
    // Log by log: wait for the group that feeds it (the other groups may still be computing),
    // fetch its {count, overflow}, patch.  The entries themselves are already in host memory.
    long long total = 0;
    bool overflow = false;
    double wait_ms = 0, patch_ms = 0;
    const bool per_group_heads = grouped && heads_ready;
    if (!per_group_heads) {
        // Something ran after the last grouped step (a mitigation / reset kernel on the handle's stream
        // may still be appending to any log), or the steps ran on the handle's stream in the first place:
        // fetch every head on the handle's stream, which is ordered after all of it (it joined every
        // group's `done` event when the steps were enqueued).  A group stream is NOT ordered after the
        // setup kernels, so a head read there could miss their entries -- which the reset of the
        // counters below would then drop for good.
        CU(cudaMemcpyAsync(s->log_head, s->log_counts, (size_t)nl * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        CU(cudaStreamSynchronize(s->stream));
    }
    for (int g = 0; g < nl && !overflow; ++g) {
        const double ta = debug ? now_ms() : 0.0;
        // heads fetched by the step itself: each rides behind its group's last kernel
        if (per_group_heads) CU(cudaEventSynchronize(s->groups[g].done));
        const unsigned long long cnt = s->log_head[2 * g];
        if ((s->log_head[2 * g + 1] & 0xFFFFFFFFull) != 0 || cnt > (unsigned long long)d.logs[g].cap) {
            overflow = true;
            break;
        }
        const double tb = debug ? now_ms() : 0.0;
        if (cnt > 0)
            apply_log(s, s->log_mapped[g], (long long)cnt, mirror, (unsigned long long)d.log_e0[g] * d.plane,
                      (unsigned long long)(d.log_e0[g + 1] - d.log_e0[g]) * d.plane, single_step);
        total += (long long)cnt;
        if (debug) {
            wait_ms += tb - ta;
            patch_ms += now_ms() - tb;
        }
    }
    CU(cudaStreamSynchronize(s->stream));  // every group has joined the handle's stream
    CU(cudaMemsetAsync(s->log_counts, 0, (size_t)d.n_logs * 2 * sizeof(unsigned long long), s->stream));
    if (overflow) {  // a log ran full: what was patched so far is consistent but incomplete
        if ((rc = download_maps(s, 0, d.E, mirror))) return rc;
        CU(cudaStreamSynchronize(s->stream));
        total = -1;
    }
    if (n_changes) *n_changes = total;
    s->mirror = mirror;
    s->full_resync = 0;
    if (debug)
        fprintf(stderr, "[sfb_sync] %d logs, %lld entries (%s): waiting for the device %.3f ms, patching %.3f ms%s\n", nl, total,
                single_step ? "one step: one-pass patch" : "ordered two-pass patch", wait_ms, patch_ms,
                overflow ? " (overflow: full download)" : "");
    return 0;
}

extern "C" int sfb_set_tracking(sfb_sim* s, int32_t enabled) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_set_tracking: null handle");
    if (!s->log_counts) return fail(SFB_ERR_STATE, "sfb_set_tracking: the handle was created without SFB_TRACK_CHANGES");
    if (s->in_step) return fail(SFB_ERR_STATE, "sfb_set_tracking: a step is half done");
    const int on = enabled != 0;
    if (on && !s->d.track) s->full_resync = 1;  // changes made while paused were not logged
    s->d.track = on;
    return 0;
}

extern "C" int sfb_get_plane(sfb_sim* s, int32_t env, int32_t plane, void* out) {
    if (!s || !out) return fail(SFB_ERR_INVALID, "sfb_get_plane: null argument");
    if (env < 0 || env >= s->d.E) return fail(SFB_ERR_INVALID, "sfb_get_plane: env %d of %d", env, s->d.E);
    if (plane < 0 || plane > SFB_PLANE_IGNITION) return fail(SFB_ERR_INVALID, "sfb_get_plane: plane %d", plane);
    if (plane == SFB_PLANE_ROS && !s->d.keep_ros)
        return fail(SFB_ERR_STATE, "sfb_get_plane: rate_of_spread is only kept with SFB_KEEP_ROS");
    if (plane == SFB_PLANE_IGNITION && !s->d.ign)
        return fail(SFB_ERR_STATE, "sfb_get_plane: the ignition plane is only kept with SFB_KEEP_IGNITION");
    int rc;
    if ((rc = use(s))) return rc;
    const size_t hw = (size_t)s->d.H * s->d.W;
    const size_t esz = plane <= SFB_PLANE_ROS ? 8 : ((plane == SFB_PLANE_AGE || plane == SFB_PLANE_IGNITION) ? 4 : 1);
    if ((rc = ensure_stage(s, hw * esz))) return rc;
    DISPATCH(s, k_get_plane, nblocks((long long)hw, 256), 256, s->d, s->parity, env, plane, s->stage);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, s->stage, hw * esz, cudaMemcpyDeviceToHost, s->stream));
    CU(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int sfb_get_status(sfb_sim* s, int32_t* status, double* elapsed, int32_t* steps) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_get_status: null handle");
    int rc;
    if ((rc = use(s))) return rc;
    return read_status(s, 0, s->d.E, status, elapsed, steps);
}

extern "C" int sfb_fire_map_device(sfb_sim* s, void** dev) {
    if (!s || !dev) return fail(SFB_ERR_INVALID, "sfb_fire_map_device: null argument");
    int rc;
    if ((rc = use(s))) return rc;
    const size_t bytes = (size_t)s->d.E * s->d.H * s->d.W;
    if (!s->obs && (rc = dmalloc(s, (char**)&s->obs, bytes))) return rc;
    if (linear_bytes(s)) {
        SFB_LAUNCH(k_get_map_v16, cap_grid(s, (long long)bytes / 16, 256), 256, 0, s->stream, (const uint4*)s->d.state, (uint4*)s->obs, (long long)bytes / 16);
        s->launches_all++;
    } else
        DISPATCH(s, k_get_map, cap_grid(s, (long long)bytes, 256), 256, s->d, 0, s->d.E, (int8_t*)s->obs);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s->stream));
    *dev = s->obs;
    return 0;
}

// ---------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------
extern "C" int sfb_get_stream(sfb_sim* s, void** stream) {
    if (!s || !stream) return fail(SFB_ERR_INVALID, "sfb_get_stream: null argument");
    *stream = (void*)s->stream;
    return 0;
}

extern "C" int sfb_get_launch_counts(sfb_sim* s, int64_t* all_kernels, int64_t* step_kernels) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_get_launch_counts: null handle");
    if (all_kernels) *all_kernels = s->launches_all;
    if (step_kernels) *step_kernels = s->launches_step;
    return 0;
}

extern "C" int sfb_set_kernel_timing(sfb_sim* s, int32_t enabled) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_set_kernel_timing: null handle");
    s->timing = enabled != 0;
    s->sweep_ms = s->rows_ms = s->eval_ms = 0;
    s->timed_steps = 0;
    return 0;
}

extern "C" int sfb_get_kernel_ms(sfb_sim* s, double* sweep_ms, double* rows_ms, double* eval_ms, int64_t* n_steps) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_get_kernel_ms: null handle");
    if (sweep_ms) *sweep_ms = s->sweep_ms;
    if (rows_ms) *rows_ms = s->rows_ms;
    if (eval_ms) *eval_ms = s->eval_ms;
    if (n_steps) *n_steps = s->timed_steps;
    s->sweep_ms = s->rows_ms = s->eval_ms = 0;
    s->timed_steps = 0;
    return 0;
}

extern "C" int sfb_get_queue_stats(sfb_sim* s, int64_t* entries, int64_t* capacity, int32_t* overflowed) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_get_queue_stats: null handle");
    int rc;
    if ((rc = use(s))) return rc;
    // the step that ran last used parity^1; k_eval leaves its counters in place until the
    // step after next resets them
    const int par = s->parity ^ 1;
    CU(cudaStreamSynchronize(s->stream));
    if (s->front_lists) {  // the watch list the next step reads; "overflowed" = the handle went dense
        unsigned long long c[4];
        CU(cudaMemcpy(c, s->list_ctr, sizeof(c), cudaMemcpyDeviceToHost));
        s->list_entries_last = (int64_t)c[s->lpar];
        if (entries) *entries = (int64_t)c[s->lpar];
        if (capacity) *capacity = s->d.wl_cap;
        if (overflowed) *overflowed = (int32_t)(c[3] & 0xFFFFFFFFull) != 0;
        return 0;
    }
    int64_t tot = 0, cap = 0;
    int32_t any_ovf = 0;
    std::vector<EnvGroup*> views;
    if (s->last_mode == 2) for (auto& gr : s->groups) views.push_back(&gr);
    else views.push_back(&s->all);
    for (EnvGroup* grp : views) {
        EnvGroup& gr = *grp;
        unsigned long long c[10];
        CU(cudaMemcpy(c, gr.counters, sizeof(c), cudaMemcpyDeviceToHost));
        tot += (int64_t)c[par];
        cap += gr.d.qcap;
        any_ovf |= reinterpret_cast<int32_t*>(c + 8)[par];
    }
    if (entries) *entries = tot;
    if (capacity) *capacity = cap;
    if (overflowed) *overflowed = any_ovf;
    return 0;
}

extern "C" int sfb_get_row_tasks(sfb_sim* s, int64_t* tasks, int64_t* capacity) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_get_row_tasks: null handle");
    int rc;
    if ((rc = use(s))) return rc;
    const int par = s->parity ^ 1;
    CU(cudaStreamSynchronize(s->stream));
    if (s->front_lists) {  // no row tasks: the step walks its watch list
        if (tasks) *tasks = 0;
        if (capacity) *capacity = 0;
        return 0;
    }
    int64_t tot = 0, cap = 0;
    std::vector<EnvGroup*> views;
    if (s->last_mode == 2) for (auto& gr : s->groups) views.push_back(&gr);
    else views.push_back(&s->all);
    for (EnvGroup* grp : views) {
        EnvGroup& gr = *grp;
        unsigned long long c[10];
        CU(cudaMemcpy(c, gr.counters, sizeof(c), cudaMemcpyDeviceToHost));
        tot += (int64_t)c[(gr.d.bits ? 10 : 4) + par];
        cap += gr.d.rows_cap;
    }
    if (tasks) *tasks = tot;
    if (capacity) *capacity = cap;
    return 0;
}

extern "C" int sfb_get_unit_stats(sfb_sim* s, int64_t* listed, int64_t* total, int32_t* mode) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_get_unit_stats: null handle");
    int rc;
    if ((rc = use(s))) return rc;
    const int par = s->parity ^ 1;
    CU(cudaStreamSynchronize(s->stream));
    int64_t tot = s->d.n_units, act = s->d.n_units;
    if (s->front_lists) {
        unsigned long long c[2];
        CU(cudaMemcpy(c, s->list_ctr, sizeof(c), cudaMemcpyDeviceToHost));
        if (listed) *listed = (int64_t)c[s->lpar];
        if (total) *total = (int64_t)s->d.E * s->d.H * s->d.W;
        if (mode) *mode = 3;
        return 0;
    }
    if (s->front_bits) {  // the listed units are the tile tasks
        act = 0;
        tot = (int64_t)s->d.E * s->d.tiles_y * s->d.tiles_x;
        std::vector<EnvGroup*> views;
        if (s->last_mode == 2) for (auto& gr : s->groups) views.push_back(&gr);
        else views.push_back(&s->all);
        for (EnvGroup* grp : views) {
            unsigned long long c[N_COUNTERS];
            CU(cudaMemcpy(c, grp->counters, sizeof(c), cudaMemcpyDeviceToHost));
            act += (int64_t)c[10 + par];  // k_eval keeps the length of the list it emptied here
        }
        if (listed) *listed = act;
        if (total) *total = tot;
        if (mode) *mode = 4;
        return 0;
    }
    if (s->unit_skip) {
        act = 0;
        std::vector<EnvGroup*> views;
        if (s->last_mode == 2) for (auto& gr : s->groups) views.push_back(&gr);
        else views.push_back(&s->all);
        for (EnvGroup* grp : views) {
            unsigned long long c[N_COUNTERS];
            CU(cudaMemcpy(c, grp->counters, sizeof(c), cudaMemcpyDeviceToHost));
            act += (int64_t)c[(s->unit_rows ? 4 : 10) + par];  // row units: the listed units are the row tasks
        }
    }
    if (listed) *listed = act;
    if (total) *total = tot;
    if (mode) *mode = !s->unit_skip ? 0 : (s->unit_rows ? 2 : 1);
    return 0;
}

extern "C" int sfb_get_front_stats(sfb_sim* s, int64_t* stats, int32_t n) {
    if (!s || !stats) return fail(SFB_ERR_INVALID, "sfb_get_front_stats: null argument");
    if (n < 0 || n > FRONT_N_STATS) return fail(SFB_ERR_INVALID, "sfb_get_front_stats: n = %d of %d", n, FRONT_N_STATS);
    if (!s->front_lists && !s->front_bits) return fail(SFB_ERR_STATE, "sfb_get_front_stats: not a list or bitboard handle");
    int rc;
    if ((rc = use(s))) return rc;
    unsigned long long c[FRONT_N_STATS];
    unsigned long long* dev = s->front_lists ? s->d.front_stats : s->list_ctr + 8;
    CU(cudaMemcpyAsync(c, dev, sizeof(c), cudaMemcpyDeviceToHost, s->stream));
    CU(cudaMemsetAsync(dev, 0, sizeof(c), s->stream));
    CU(cudaStreamSynchronize(s->stream));
    for (int k = 0; k < n; ++k) stats[k] = (int64_t)c[k];
    return 0;
}

extern "C" int sfb_debug_stall(sfb_sim* s, int32_t microseconds) {
    if (!s) return fail(SFB_ERR_INVALID, "sfb_debug_stall: null handle");
    if (microseconds < 0 || microseconds > 1000000) return fail(SFB_ERR_INVALID, "sfb_debug_stall: %d us", microseconds);
    int rc;
    if ((rc = use(s))) return rc;
    SFB_LAUNCH(k_stall, 1, 1, 0, s->stream, (long long)microseconds * 1000);
    s->launches_all++;
    CU(cudaGetLastError());
    return 0;
}

extern "C" int sfb_device_bytes(sfb_sim* s, int64_t* bytes) {
    if (!s || !bytes) return fail(SFB_ERR_INVALID, "sfb_device_bytes: null argument");
    *bytes = s->dev_bytes;
    return 0;
}

extern "C" int sfb_rate_of_spread(int32_t device, const int8_t* dir, const float* rec, const float* particle,
                                  int64_t n, double* out) {
    if (!dir || !rec || !particle || !out || n < 0) return fail(SFB_ERR_INVALID, "sfb_rate_of_spread: bad argument");
    if (n == 0) return 0;
    CU(cudaSetDevice(device));
    int8_t* d_dir = nullptr;
    float* d_rec = nullptr;
    double* d_out = nullptr;
    auto body = [&]() -> int {
        CU(cudaMalloc((void**)&d_dir, (size_t)n));
        CU(cudaMalloc((void**)&d_rec, (size_t)n * 8 * sizeof(float)));
        CU(cudaMalloc((void**)&d_out, (size_t)n * sizeof(double)));
        CU(cudaMemcpy(d_dir, dir, (size_t)n, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_rec, rec, (size_t)n * 8 * sizeof(float), cudaMemcpyHostToDevice));
        SfbParticle fp{particle[0], particle[1], particle[2], particle[3], particle[4]};
        SFB_LAUNCH(k_rate_of_spread, nblocks(n, 128), 128, 0, nullptr, d_dir, d_rec, fp, (long long)n, d_out);
        CU(cudaGetLastError());
        CU(cudaMemcpy(out, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
        return 0;
    };
    const int rc = body();
    cudaFree(d_dir);
    cudaFree(d_rec);
    cudaFree(d_out);
    return rc;
}
