// List-driven step ("watch list" front end): the default for every handle that is not a slab.
//
// The reference's update() (simfire/game/managers/fire.py:616-719) walks its list of Fire sprites;
// its cost is proportional to the fire fronts, not to the grid.  This is that algorithm in its
// data-parallel form.  The handle keeps ONE list of the cells that can change in a step:
//
//   * cells that carry a Fire sprite (they age and burn out: fire.py:116-161, :633),
//   * ignitable cells next to a live sprite (the candidates of fire.py:163-234 / :451-517),
//   * with attenuation, every control-line cell (fire.py:271-278 subtracts from all of them),
//
// each at most once: `listed` holds one bit per cell, set by an atomic test-and-set when a cell
// joins the list and cleared when it leaves.  One step is ONE kernel (k_front), one thread per
// entry, the entry's thread being the only writer of that cell's state byte and burn value:
//
//   sprite role  code != 0: duration = (t - 1 - ign) mod M >= max_fire_duration -> BURNED (prune),
//                else the env has a live sprite (fire.py:637);
//   watch role   ignitable: the eight neighbours' bytes are read, the source whose pair the
//                reference writes last wins (smallest duration, then S-E, S, S-W, E, W, N-E, N, N-W:
//                fire.py:704-705 + sprite-list order), rate of spread of (cell, direction) from the
//                per-cell table, attenuation, float64 burn accumulation, ignition on burn > pixel_scale;
//   adds         a sprite of duration 0 (ignited by the previous step, or the initial fire) test-and-
//                sets the bits of its ignitable neighbours; those it wins are examined in the SAME
//                step by the threads of the block (a shared-memory task list keeps that balanced) and
//                join the list.  Cells dropped this step have no live source, cells added this step
//                have one (the new sprite): the two sets are disjoint, so clearing and setting bits
//                never race.
//
// Races between threads are benign by construction (same argument as in k_rows): a neighbour's byte
// is either its value before this step or after it; a sprite pruned this step is "not a live source"
// in both, and a cell ignited this step carries code(t), whose duration (t - 1 - t) mod M = M - 1 is
// never below max_fire_duration.
//
// Survivors are staged in shared memory and appended to the other list buffer with one global atomic
// per flush (the list tail is a single address).  The last block to finish (atomic ticket) advances
// the per-env clocks (fire.py:633-652, :717).  With attenuation a second kernel (k_tail) applies the
// deferred subtraction to control-line cells no fire touches, which depends on a whole-env flag of
// the same step (fire.py:651-652), and then advances the clocks.
//
// If an append ever finds the list full, a sticky flag turns the handle to the dense form of the
// same per-cell routine (one thread per cell, no list): slow, but identical results.
#pragma once
#include "sfb_kernels.cuh"

namespace sfb {

// ---- entries ----------------------------------------------------------------------------
constexpr int LE_BITS = 20;  // x and y: sfb_create rejects H, W >= 2^20; env: 22 bits; deferred status: 2 bits
__device__ __forceinline__ unsigned long long le_make(int env, int y, int x) {
    return (unsigned long long)(unsigned)x | ((unsigned long long)(unsigned)y << LE_BITS) |
           ((unsigned long long)(unsigned)env << (2 * LE_BITS));
}
__device__ __forceinline__ int le_x(unsigned long long e) { return (int)(e & ((1u << LE_BITS) - 1u)); }
__device__ __forceinline__ int le_y(unsigned long long e) { return (int)((e >> LE_BITS) & ((1u << LE_BITS) - 1u)); }
__device__ __forceinline__ int le_env(unsigned long long e) { return (int)((e >> (2 * LE_BITS)) & 0x3FFFFFu); }
__device__ __forceinline__ int le_defer(unsigned long long e) { return (int)(e >> 62); }  // 0, or internal status - 3

// neighbours in the order in which the reference's last write wins (rank 0 first): offset of the
// SOURCE relative to the destination, and the index of that pair's direction in fire.py:211-221
__device__ __forceinline__ void rank_offset(int r, int& dy, int& dx) {
    dy = r < 3 ? 1 : (r < 5 ? 0 : -1);
    dx = (0x01202012u >> (4 * r)) & 0xF;  // ranks 0..7: +1, 0, -1, +1, -1, +1, 0, -1  (stored + 1)
    dx -= 1;
}
constexpr uint32_t RANK_TO_DIR = 0x12304765u;  // nibble r = direction index of rank r
constexpr uint32_t ORTHO_RANKS = 0x5Au;        // ranks 1 (S), 3 (E), 4 (W), 6 (N): the 4-neighbour variant (fire.py:223-228)

__device__ __forceinline__ bool listed_test_and_set(const DevParams& p, long long idx) {
    const uint32_t bit = 1u << (idx & 31);
    return (atomicOr(p.listed + (idx >> 5), bit) & bit) == 0;  // true: the caller put it on the list
}
// the same as a predicated instruction (no branch around it): a lane's eight test-and-sets are all in
// flight before the first result is looked at.  Returns the old word, all ones ("listed") if !pred.
__device__ __forceinline__ uint32_t listed_or_pred(const DevParams& p, long long idx, bool pred) {
    const uint32_t bit = 1u << (idx & 31);
#ifndef SFB_EMU
    uint32_t old = 0xFFFFFFFFu;
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "setp.ne.u32 q, %3, 0;\n\t"
        "@q atom.global.or.b32 %0, [%1], %2;\n\t}"
        : "+r"(old)
        : "l"(p.listed + (idx >> 5)), "r"(bit), "r"((uint32_t)pred)
        : "memory");
    return old;
#else
    return pred ? atomicOr(p.listed + (idx >> 5), bit) : 0xFFFFFFFFu;
#endif
}
__device__ __forceinline__ void listed_clear(const DevParams& p, long long idx) {
    atomicAnd(p.listed + (idx >> 5), ~(1u << (idx & 31)));
}

// rate of spread of the pair that travels in direction `dir` into static cell `sc` (ft/min, float64):
// from the per-cell table k_derive_static filled (every input of rothermel.py:4-136 is static per
// cell and direction), else evaluated from the derived fuel terms
template <bool RTAB>
__device__ __forceinline__ double pair_rate(const DevParams& p, long long sc, int dir) {
    if (RTAB) return __ldg(p.rtab + sc * 8 + dir);
    const float4* rp = reinterpret_cast<const float4*>(p.drv + sc);
    const float4 t0 = __ldg(rp), t1 = __ldg(rp + 1), e = __ldg(rp + 2);
    const SfbFuelTerms t = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    return sfb_spread_from_terms(dir, t, e.x, e.y, e.z, e.w);
}

// L2 prefetch (a hint: no register, nobody waits for it).  k_front issues these for the cells it will
// examine two tiles from now and for the cells it has just put on the list: the dependent gathers of
// the examination then find their sectors (and their page translations) on the way or already there.
__device__ __forceinline__ void prefetch_l2(const void* ptr) {
#if !defined(SFB_EMU) && !defined(SFB_NO_PREFETCH)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}
template <typename CellT, bool RTAB>
__device__ __forceinline__ void prefetch_cell(const DevParams& p, const int par, const int env, const int y, const int x) {
    const long long cell = (long long)y * p.pitch + x;
    const long long idx = (long long)env * p.plane + cell;
    const CellT* st = reinterpret_cast<const CellT*>(p.state) + idx;
    prefetch_l2(st);
    if (y > 0) prefetch_l2(st - p.pitch);
    if (y + 1 < p.H) prefetch_l2(st + p.pitch);
    prefetch_l2(p.burn + idx);
    if (RTAB) {
        const double* rt = p.rtab + (p.shared_static ? cell : idx) * 8;
        prefetch_l2(rt);
        prefetch_l2(rt + 4);
    }
}

struct ExamOut {
    bool keep;       // the cell stays on the list
    int defer;       // control-line cell no fire touches: internal status (4..6), else 0 (fire.py:271-278, :651)
    uint32_t push;   // bit r: the neighbour at rank_offset(r) is ignitable and has to be on the list (this cell
                     // is a sprite of duration 0)
    int log;         // BurnStatus to append to the change log, or -1
    bool ros_have;   // the cell was a candidate; SFB_KEEP_ROS: its rate_of_spread entry of this step
    double ros;
    bool nbr;        // the eight neighbours were read
    bool live;       // the cell carries a sprite that survives this call's pruning (fire.py:637)
};
// statistics of the last k_front launches (sfb_get_front_stats): cells examined, candidates evaluated,
// cells ignited, sprites pruned, cells that joined the list, list entries read, cells whose neighbours were read
constexpr int FRONT_N_STATS = 7;

// One cell of one env, one update() call.  LISTS: the cell is on the watch list (bits and pushes are
// maintained); otherwise the dense form (every cell is visited, nothing to maintain).
template <typename CellT, bool LISTS, bool RTAB>
__device__ __forceinline__ ExamOut examine_cell(const DevParams& p, const int par, const int env, const int y, const int x) {
    using C = Cell<CellT>;
    ExamOut o;
    o.keep = false;
    o.defer = 0;
    o.push = 0;
    o.log = -1;
    o.ros_have = false;
    o.ros = 0.0;
    o.nbr = false;
    o.live = false;
    // {t, running, time_quit, any_live}: one 16-byte load (any_live may be changing under us: not used here)
    const int4 m4 = *reinterpret_cast<const int4*>(p.meta + (long long)par * p.meta_stride + env);
    const long long idx = (long long)env * p.plane + (long long)y * p.pitch + x;
    CellT* const state = reinterpret_cast<CellT*>(p.state);
    const int c = state[idx];
    const int t = m4.x;
    if (!m4.y) return o;  // frozen until sfb_reset (which purges the env's entries and bits)
    const bool spread = !m4.z;  // fire.py:641-643: the call prunes, then returns QUIT
    int s = c & 7;
    const int code = c >> 3;
    const int tm1 = (t - 1) % C::M;
    bool live = false, age0 = false;
    if (code != 0) {
        const int a = sprite_age<CellT>(code, tm1);
        if (a >= p.max_dur) {  // fire.py:116-161
            state[idx] = (CellT)ST_BURNED;
            s = ST_BURNED;
            o.log = 2;
        } else {
            live = true;
            age0 = a == 0;
            o.live = true;  // fire.py:637 (the caller raises the env's flag)
        }
    }
    const bool watch = spread && ignitable(s);
    o.keep = live;
    if (!(watch || (LISTS && age0 && spread))) return o;

    // the eight neighbours (BURNED stands in for cells outside the grid: no source, not ignitable); the
    // cell's burn value is fetched alongside them (it is needed one dependent load later otherwise)
    double b = 0.0;
    if (watch) b = p.burn[idx];
    o.nbr = true;
    const uint32_t ranks = p.diagonal ? 0xFFu : ORTHO_RANKS;
    // all eight loads are issued before the first byte is looked at (predicated loads, no branches in
    // between): the kernel is bound by the latency of its dependent gathers
    int nb[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        int dy, dx;
        rank_offset(r, dy, dx);
        const bool inb = ((ranks >> r) & 1u) && (unsigned)(y + dy) < (unsigned)p.H && (unsigned)(x + dx) < (unsigned)p.W;
        nb[r] = ST_BURNED;
        if (inb) nb[r] = state[idx + (long long)dy * p.pitch + dx];
    }
    int best = 1 << 20;
    uint32_t push = 0;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int ncode = nb[r] >> 3;
        const int a = sprite_age<CellT>(ncode, tm1);  // a cell ignited by this very step has duration M - 1
        const int key = (ncode != 0 && a < p.max_dur) ? a * 8 + r : (1 << 20);
        best = min(best, key);
        push |= ignitable(nb[r] & 7) ? 1u << r : 0u;
    }
    if (LISTS && age0 && spread) o.push = push;
    if (!watch) return o;

    if (best < (1 << 20)) {  // a candidate (fire.py:163-234; the caller raises the env's flag, fire.py:651)
        const int dir = (RANK_TO_DIR >> ((best & 7) * 4)) & 0xF;
        double ros = pair_rate<RTAB>(p, p.shared_static ? idx - (long long)env * p.plane : idx, dir) * p.dt;  // fire.py:696
        if (s & ST_LINE_BIT) ros = p.attenuate ? ros - line_attenuation(s) : 0.0;                   // fire.py:271-282
        o.ros_have = true;
        o.ros = ros;
        if (ros != 0.0) {  // burn + 0 == burn: skip the store
            b += ros;      // fire.py:710
            p.burn[idx] = b;
        }
        if (b > p.ps) {  // fire.py:568 (strict)
            if (live) {
                // a control line drawn over a burning cell re-ignites while its first sprite is still a source
                // for its neighbours in THIS step (the newer sprite replaces the older one, as in the oracle):
                // the byte is rewritten when every cell has been examined (k_front's last block)
                const unsigned int slot = atomicAdd(p.late_count, 1u);
                if (slot < (unsigned int)p.late_cap) p.late[slot] = (unsigned long long)idx;
                else *p.broken = 2;  // reported by sfb_synchronize
            } else {
                const int ncode = 1 + (t % C::M);
                state[idx] = (CellT)(ST_BURNING | (ncode << 3));  // fire.py:571-587
                if (p.ign) p.ign[idx] = t;
            }
            o.log = 1;
        }
        o.keep = true;  // still a candidate, or a sprite from the next step on
    } else if ((s & ST_LINE_BIT) && p.attenuate) {
        o.defer = s;    // attenuated only if the env gets past the early return (fire.py:651-652): k_tail
        o.keep = true;
    }
    return o;
}

// ---- k_front ----------------------------------------------------------------------------
// Warps work on their own: each takes tiles of 32 entries (round-robin over the grid), stages what
// survives in its own shared-memory buffer and flushes it with one global atomic when the next tile
// might not fit.  No block-wide barrier in the loop: the kernel is bound by the latency of its
// dependent gathers (entry -> state byte -> neighbours / burn -> rate), so nothing may wait for the
// slowest of 256 threads.
constexpr int FRONT_THREADS = 256;
constexpr int FRONT_WARPS = FRONT_THREADS / 32;
constexpr int FRONT_TASK_CAP = 32 * 8;                                // a tile's sprites of duration 0 push at most 8 cells each
constexpr int FRONT_OUT_CAP = 512;                                    // staged survivors per warp
constexpr int FRONT_TILE_OUT = 32 + FRONT_TASK_CAP;                   // entries one tile can keep
static_assert(FRONT_OUT_CAP >= FRONT_TILE_OUT, "a tile must fit the stage");
constexpr int FRONT_SMEM = FRONT_WARPS * (FRONT_OUT_CAP + FRONT_TASK_CAP) * 8;

// the per-env clock of the next step (fire.py:633-652, :717); reads what this step finalised
__device__ __forceinline__ void advance_clock(const DevParams& p, const int par, const int env) {
    const volatile EnvMeta* cp = p.meta + (long long)par * p.meta_stride + env;
    EnvMeta cur;
    cur.t = cp->t;
    cur.running = cp->running;
    cur.time_quit = cp->time_quit;
    cur.any_live = cp->any_live;
    cur.any_cand = cp->any_cand;
    cur.pad = 0;
    cur.elapsed = cp->elapsed;
    EnvMeta nxt = cur;
    if (cur.running) {
        if (!cur.any_live) nxt.running = 0;            // fire.py:637
        else if (cur.time_quit) nxt.running = 0;       // fire.py:641-643
        else if (cur.any_cand) nxt.elapsed = cur.elapsed + p.dt;  // fire.py:717 (skipped by :651)
        nxt.t = cur.t + 1;
    }
    nxt.any_live = 0;
    nxt.any_cand = 0;
    nxt.time_quit = p.has_max_time && (p.dt > p.max_time || nxt.elapsed > p.max_time);
    p.meta[(long long)(par ^ 1) * p.meta_stride + env] = nxt;
}

template <typename CellT, bool RTAB>
__global__ void __launch_bounds__(FRONT_THREADS, 4) k_front(const DevParams p, const int par, const int lpar) {
    SFB_DYNAMIC_SMEM(smem_raw);
    __shared__ unsigned int s_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned long long* const w_out = reinterpret_cast<unsigned long long*>(smem_raw) + warp * (FRONT_OUT_CAP + FRONT_TASK_CAP);
    unsigned long long* const w_task = w_out + FRONT_OUT_CAP;
    unsigned long long* const wout = p.wl[lpar ^ 1];
    const uint32_t lt = (1u << lane) - 1u;
    int out_cnt = 0;  // staged survivors of this warp (warp-uniform)

    auto stage = [&](bool keep, unsigned long long e) {
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) w_out[out_cnt + __popc(m & lt)] = e;
        out_cnt += __popc(m);
    };
    auto flush = [&]() {  // one global atomic per flush: the list tail is a single address
        if (out_cnt == 0) return;
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.wl_count + (lpar ^ 1), (unsigned long long)out_cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < out_cnt; i += 32) {
            if (base + i < (unsigned long long)p.wl_cap) wout[base + i] = w_out[i];
            else *p.broken = 1;  // the list is full: dense form from the next step on
        }
        out_cnt = 0;
        __syncwarp();
    };
    // what every examined cell leaves behind besides its list entry (warp-wide calls)
    int n_cand = 0, n_ign = 0, n_prune = 0, n_exam = 0, n_task_done = 0, n_nbr = 0;
    auto side_effects = [&](const ExamOut& o, int env, int y, int x) {
        const long long idx = (long long)env * p.plane + (long long)y * p.pitch + x;
        n_cand += o.ros_have;
        n_nbr += o.nbr;
        n_ign += o.log == 1;
        n_prune += o.log == 2;
        // the env-wide flags of fire.py:637 / :651: entries of one env come in runs, so a lane only stores
        // a flag that the lane before it does not already raise for the same env
        const int key = (env << 2) | (o.live ? 1 : 0) | (o.ros_have ? 2 : 0);
        int prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (lane == 0) prev = -4;
        const bool same = (prev >> 2) == env;
        EnvMeta* const mp = p.meta + (long long)par * p.meta_stride + env;
        if (o.live && !(same && (prev & 1))) mp->any_live = 1;
        if (o.ros_have && !(same && (prev & 2))) mp->any_cand = 1;
        if (p.track) log_append(p, o.log >= 0, idx, o.log);
        if (p.keep_ros) {
            const uint32_t m = __ballot_sync(0xffffffffu, o.ros_have);
            if (m) {
                unsigned long long base = 0;
                if (lane == __ffs(m) - 1) base = atomicAdd(p.ros_count, (unsigned long long)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (o.ros_have) {
                    const unsigned long long slot = base + __popc(m & lt);
                    if (slot < (unsigned long long)p.ros_cap) {
                        p.ros_items[2 * slot] = (unsigned long long)idx;
                        p.ros_items[2 * slot + 1] = (unsigned long long)__double_as_longlong(o.ros);
                    }
                }
            }
        }
    };
    auto empty_exam = []() {
        ExamOut o;
        o.keep = false;
        o.defer = 0;
        o.push = 0;
        o.log = -1;
        o.ros_have = false;
        o.ros = 0.0;
        o.nbr = false;
        o.live = false;
        return o;
    };

    if (*p.dense_now) {  // (latched between launches: every block of a launch takes the same form)
        // dense form: every cell of every env; nothing to maintain.  (Cells ignited by this pass carry
        // code(t): no source for anybody, not ignitable, so the order of the threads does not matter.)
        const long long total = (long long)p.E * p.plane;
        const long long first = (long long)blockIdx.x * FRONT_THREADS + tid - lane;
        for (long long base = first; base < total; base += (long long)gridDim.x * FRONT_THREADS) {
            const long long i = base + lane;
            ExamOut o = empty_exam();
            int env = 0, y = 0, x = 0;
            if (i < total) {
                env = (int)(i / p.plane);
                const long long cell = i - (long long)env * p.plane;
                y = (int)(cell / p.pitch);
                x = (int)(cell - (long long)y * p.pitch);
                if (x < p.W) {
                    o = examine_cell<CellT, false, RTAB>(p, par, env, y, x);
                    ++n_exam;
                }
            }
            side_effects(o, env, y, x);
        }
    } else {
        const unsigned long long n_in = min(p.wl_count[lpar], (unsigned long long)p.wl_cap);
        const unsigned long long* const win = p.wl[lpar];
        if (p.front_stats && blockIdx.x == 0 && tid == 0) atomicAdd(p.front_stats + 5, n_in);
        const unsigned long long n_warps = (unsigned long long)gridDim.x * FRONT_WARPS;
        const unsigned long long tile0 = (unsigned long long)blockIdx.x * FRONT_WARPS + warp;
        // entries are fetched two tiles ahead and their cells prefetched one tile ahead of their examination
        constexpr unsigned long long NONE = ~0ull;
        unsigned long long e_cur = tile0 * 32 + lane < n_in ? win[tile0 * 32 + lane] : NONE;
        unsigned long long e_nxt = (tile0 + n_warps) * 32 + lane < n_in ? win[(tile0 + n_warps) * 32 + lane] : NONE;
        if (e_cur != NONE) prefetch_cell<CellT, RTAB>(p, par, le_env(e_cur), le_y(e_cur), le_x(e_cur));
        for (unsigned long long tile = tile0; tile * 32 < n_in; tile += n_warps) {
            const unsigned long long i2 = (tile + 2 * n_warps) * 32 + lane;
            const unsigned long long e_nn = i2 < n_in ? win[i2] : NONE;
            if (e_nxt != NONE) prefetch_cell<CellT, RTAB>(p, par, le_env(e_nxt), le_y(e_nxt), le_x(e_nxt));
            // phase 1: this lane's own entry
            ExamOut o = empty_exam();
            int env = 0, y = 0, x = 0;
            if (e_cur != NONE) {
                const unsigned long long e = e_cur;
                env = le_env(e);
                y = le_y(e);
                x = le_x(e);
                o = examine_cell<CellT, true, RTAB>(p, par, env, y, x);
                ++n_exam;
                if (!o.keep) listed_clear(p, (long long)env * p.plane + (long long)y * p.pitch + x);
            }
            stage(o.keep, le_make(env, y, x) | ((unsigned long long)(o.defer ? o.defer - 3 : 0) << 62));
            side_effects(o, env, y, x);
            // ... and the neighbours its new sprite brings onto the list: tasks for the whole warp
            int n_task = 0;
            if (__any_sync(0xffffffffu, o.push != 0)) {
                uint32_t won = 0, old[8];  // all test-and-sets are issued before the first result is looked at
                const long long idx0 = (long long)env * p.plane + (long long)y * p.pitch + x;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    int dy, dx;
                    rank_offset(r, dy, dx);
                    old[r] = listed_or_pred(p, idx0 + (long long)dy * p.pitch + dx, (o.push >> r) & 1u);
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    int dy, dx;
                    rank_offset(r, dy, dx);
                    const uint32_t bit = 1u << ((idx0 + (long long)dy * p.pitch + dx) & 31);
                    won |= (old[r] & bit) ? 0u : 1u << r;
                }
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    int dy, dx;
                    rank_offset(r, dy, dx);
                    const bool mine = (won >> r) & 1u;
                    const uint32_t m = __ballot_sync(0xffffffffu, mine);
                    if (mine) {
                        w_task[n_task + __popc(m & lt)] = le_make(env, y + dy, x + dx);
                        prefetch_cell<CellT, RTAB>(p, par, env, y + dy, x + dx);
                    }
                    n_task += __popc(m);
                }
            }
            __syncwarp();
            // phase 2: the cells that joined the list are examined by whichever lane is free
            for (int base = 0; base < n_task; base += 32) {
                const int j = base + lane;
                ExamOut q = empty_exam();
                int e2 = 0, y2 = 0, x2 = 0;
                if (j < n_task) {
                    const unsigned long long e = w_task[j];
                    e2 = le_env(e);
                    y2 = le_y(e);
                    x2 = le_x(e);
                    q = examine_cell<CellT, true, RTAB>(p, par, e2, y2, x2);
                    ++n_exam;
                    ++n_task_done;
                    if (!q.keep) listed_clear(p, (long long)e2 * p.plane + (long long)y2 * p.pitch + x2);
                }
                stage(q.keep, le_make(e2, y2, x2) | ((unsigned long long)(q.defer ? q.defer - 3 : 0) << 62));
                side_effects(q, e2, y2, x2);
            }
            __syncwarp();
            if (out_cnt > FRONT_OUT_CAP - FRONT_TILE_OUT) flush();
            e_cur = e_nxt;
            e_nxt = e_nn;
        }
        flush();
    }

    if (p.front_stats) {  // one atomic per warp and counter that has something
        const int v[7] = {n_exam, n_cand, n_ign, n_prune, n_task_done, 0, n_nbr};
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            int w = v[k];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) w += __shfl_down_sync(0xffffffffu, w, d);
            if (lane == 0 && w) atomicAdd(p.front_stats + k, (unsigned long long)w);
        }
    }
    // the last block to get here closes the step
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(p.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) {
        *p.ticket = 0;
        p.wl_count[lpar] = 0;  // consumed
        if (*reinterpret_cast<volatile int32_t*>(p.broken)) *p.dense_now = 1;  // an append of this step found the list full
    }
    {   // re-ignitions of cells that were still sources during the step (rare: see examine_cell)
        const unsigned int n_late = min(*reinterpret_cast<volatile unsigned int*>(p.late_count), (unsigned int)p.late_cap);
        for (unsigned int i = tid; i < n_late; i += FRONT_THREADS) {
            const long long idx = (long long)reinterpret_cast<volatile unsigned long long*>(p.late)[i];
            const int env = (int)(idx / p.plane);
            const int t = p.meta[(long long)par * p.meta_stride + env].t;
            reinterpret_cast<CellT*>(p.state)[idx] = (CellT)(ST_BURNING | ((1 + (t % Cell<CellT>::M)) << 3));
            if (p.ign) p.ign[idx] = t;
        }
        __syncthreads();
        if (tid == 0) *p.late_count = 0;
    }
    if (!p.attenuate)  // otherwise k_tail does it, after the deferred control-line items
        for (int env = tid; env < p.E; env += FRONT_THREADS) advance_clock(p, par, env);
}

// ---- k_tail (attenuation only) ------------------------------------------------------------
// Control-line cells that are no candidates get ros = -attenuation like every other line cell of
// the map, but only if the env got past the "no new locations" early return (fire.py:651-652,
// :271-278): their entries were kept with the status in the top bits.  Then the clocks.
template <typename CellT>
__global__ void __launch_bounds__(256) k_tail(const DevParams p, const int par, const int lpar_out) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;
    if (*p.broken) {
        // dense form: recompute which line cells were no candidates.  Sources are judged with this
        // step's clock: sprites pruned by k_front were not live, cells it ignited have duration M - 1.
        using C = Cell<CellT>;
        const CellT* state = reinterpret_cast<const CellT*>(p.state);
        const long long total = (long long)p.E * p.plane;
        for (long long i = gid; i < total; i += gstride) {
            const int env = (int)(i / p.plane);
            const EnvMeta m = p.meta[(long long)par * p.meta_stride + env];
            if (!m.running || m.time_quit || !m.any_cand) continue;
            const long long cell = i - (long long)env * p.plane;
            const int y = (int)(cell / p.pitch), x = (int)(cell - (long long)y * p.pitch);
            if (x >= p.W) continue;
            const int s = state[i] & 7;
            if (!(s & ST_LINE_BIT)) continue;
            const int tm1 = (m.t - 1) % C::M;
            const uint32_t ranks = p.diagonal ? 0xFFu : ORTHO_RANKS;
            bool cand = false;
            for (int r = 0; r < 8; ++r) {
                if (!((ranks >> r) & 1u)) continue;
                int dy, dx;
                rank_offset(r, dy, dx);
                const int yy = y + dy, xx = x + dx;
                if ((unsigned)yy >= (unsigned)p.H || (unsigned)xx >= (unsigned)p.W) continue;
                const int ncode = (int)state[i + (long long)dy * p.pitch + dx] >> 3;
                if (ncode != 0 && sprite_age<CellT>(ncode, tm1) < p.max_dur) cand = true;
            }
            if (cand) continue;  // k_front handled it (and it is still a line cell: it did not ignite)
            const double ros = 0.0 - line_attenuation(s);
            if (p.keep_ros) p.ros[i] = ros;
            p.burn[i] += ros;
        }
    } else {
        const long long n = (long long)min(p.wl_count[lpar_out], (unsigned long long)p.wl_cap);
        const unsigned long long* const w = p.wl[lpar_out];
        for (long long i = gid; i < n; i += gstride) {
            const unsigned long long e = w[i];
            const int d = le_defer(e);
            if (!d) continue;
            const int env = le_env(e);
            if (!p.meta[(long long)par * p.meta_stride + env].any_cand) continue;
            const long long idx = (long long)env * p.plane + (long long)le_y(e) * p.pitch + le_x(e);
            const double ros = 0.0 - line_attenuation(d + 3);
            if (p.keep_ros) p.ros[idx] = ros;
            p.burn[idx] += ros;  // fire.py:710
        }
    }
    for (long long env = gid; env < p.E; env += gstride) advance_clock(p, par, (int)env);
}

// SFB_KEEP_ROS: the candidates' rate_of_spread values of this step, written after k_clear_ros has
// rebuilt the plane from zeros (fire.py:703-708)
__global__ void k_ros_apply(const DevParams p) {
    const long long n = (long long)min(*p.ros_count, (unsigned long long)p.ros_cap);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p.ros[(long long)p.ros_items[2 * i]] = __longlong_as_double((long long)p.ros_items[2 * i + 1]);
}

// ---- list maintenance between steps (not on the per-step path) -------------------------------
// append one cell (setup kernels: a few entries per call)
__device__ __forceinline__ void list_append(const DevParams& p, int lpar, int env, int y, int x) {
    const unsigned long long slot = atomicAdd(p.wl_count + lpar, 1ULL);
    if (slot < (unsigned long long)p.wl_cap) p.wl[lpar][slot] = le_make(env, y, x);
    else *p.broken = *p.dense_now = 1;  // between steps: the next step already runs in the dense form
}

// does (env, y, x) have a neighbour that carries a sprite code?
template <typename CellT>
__device__ __forceinline__ bool has_coded_neighbour(const DevParams& p, int env, int y, int x) {
    const CellT* st = reinterpret_cast<const CellT*>(p.state) + (long long)env * p.plane;
    for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
            if ((dy == 0 && dx == 0) || (!p.diagonal && dy != 0 && dx != 0)) continue;
            const int yy = y + dy, xx = x + dx;
            if ((unsigned)yy >= (unsigned)p.H || (unsigned)xx >= (unsigned)p.W) continue;
            if (((int)st[(long long)yy * p.pitch + xx] >> 3) != 0) return true;
        }
    return false;
}

// should a cell whose byte is `c` be on the list?  (sprite | ignitable next to a sprite | attenuated line)
template <typename CellT>
__device__ __forceinline__ bool belongs_on_list(const DevParams& p, int env, int y, int x, int c) {
    const int s = c & 7;
    if ((c >> 3) != 0) return true;
    if (!ignitable(s)) return false;
    if ((s & ST_LINE_BIT) && p.attenuate) return true;
    return has_coded_neighbour<CellT>(p, env, y, x);
}

// drop the entries of the marked envs: wl[lpar] -> wl[lpar ^ 1] (the host flips lpar afterwards)
__global__ void k_list_purge(const DevParams p, const int lpar, const uint8_t* env_mark) {
    const int lane = threadIdx.x & 31;
    const long long n = (long long)min(p.wl_count[lpar], (unsigned long long)p.wl_cap);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += stride) {
        const long long i = base + lane;
        unsigned long long e = 0;
        bool keep = false;
        if (i < n) {
            e = p.wl[lpar][i];
            keep = env_mark[le_env(e)] == 0;
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (!m) continue;
        unsigned long long slot = 0;
        if (lane == __ffs(m) - 1) slot = atomicAdd(p.wl_count + (lpar ^ 1), (unsigned long long)__popc(m));
        slot = __shfl_sync(0xffffffffu, slot, __ffs(m) - 1);
        if (keep) p.wl[lpar ^ 1][slot + __popc(m & ((1u << lane) - 1))] = e;
    }
}

// clear the `listed` bits of n envs (a device list, or [env0, env0 + n)); 16 cells per thread, plane % 16 == 0
__global__ void k_list_clear_bits(const DevParams p, const int32_t* envs, const int env0, const int n) {
    const long long plane16 = p.plane / 16, total = (long long)n * plane16;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / plane16);
        const long long env = envs ? envs[k] : env0 + k;
        const long long bit0 = (env * plane16 + (i - (long long)k * plane16)) * 16;
        atomicAnd(p.listed + (bit0 >> 5), ~(0xFFFFu << (bit0 & 31)));
    }
}

// put every cell of envs [env0, env0 + n) that belongs on the list onto wl[lpar] (after a purge and a
// clear of their bits): wholesale map replacement (sfb_set_fire_map, sfb_update)
template <typename CellT>
__global__ void k_list_rebuild(const DevParams p, const int lpar, const int env0, const int n) {
    const long long total = (long long)n * p.plane;
    const CellT* state = reinterpret_cast<const CellT*>(p.state) + (long long)env0 * p.plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / p.plane);
        const long long cell = i - (long long)k * p.plane;
        const int y = (int)(cell / p.pitch), x = (int)(cell - (long long)y * p.pitch);
        if (x >= p.W) continue;
        if (belongs_on_list<CellT>(p, env0 + k, y, x, state[i]) && listed_test_and_set(p, (long long)env0 * p.plane + i))
            list_append(p, lpar, env0 + k, y, x);
    }
}

}  // namespace sfb
