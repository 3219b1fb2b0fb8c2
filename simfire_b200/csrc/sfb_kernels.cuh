// Device side of the B200 fire-spread stepper: data layout + the two hot-path kernels.
//
// One timestep of RothermelFireManager.update (simfire/game/managers/fire.py:616-719) is
//
//   k_sweep   dense pass over the packed per-cell state (1 B/cell): prune expired sprites
//             (fire.py:116-161), find every ignitable cell that has a burning neighbour
//             and the neighbour whose pair the reference writes last (fire.py:163-234,
//             :704-705), push (cell, direction) work items to a queue, raise the per-env
//             flags the reference's early returns depend on (fire.py:637, :651).
//   k_eval    one thread per work item: Rothermel rate of spread of the destination cell
//             (rothermel.py:4-136), control-line attenuation (fire.py:236-284), burn
//             accumulation in float64 (fire.py:710), ignition on burn > pixel_scale
//             (fire.py:566-587); the first E threads also advance the per-env clock
//             (fire.py:633, :641-643, :717).
//
// HBM layout (per handle; E envs of H x W cells, rows padded to `pitch` cells):
//   state  CellT [E][H][pitch]   bits 0-2 internal status, bits 3.. sprite code (below)
//   burn   f64   [E][H][pitch]   burn_amounts
//   stat   32 B  [E or 1][H][pitch]  {w_0, delta, M_x, sigma, U, U_dir, slope_mag,
//                                slope_dir} float32 -- one 32-byte sector per cell: this
//                                record is only ever GATHERED (at work items), never swept
//   ros    f64   [E][H][pitch]   only with SFB_KEEP_ROS
//
// Internal status (bits 0-2): 0 UNBURNED, 1 BURNING, 2 BURNED, 4 FIRELINE, 5 SCRATCHLINE,
// 6 WETLINE -- BurnStatus with bit 2 meaning "control line", so that one AND per 32-bit
// word tells whether a run of cells can possibly need work this step.
// Sprite code (bits 3..): 0 = no Fire sprite on the cell; otherwise 1 + (ign mod M) where
// ign is the update() call that created the sprite (0 = initial fire) and M = 31 (8-bit
// cells) or 8191 (16-bit cells).  The duration the reference keeps per sprite
// (fire.py:633) is recovered as (t - 1 - ign) mod M, so the code never has to be
// rewritten while the sprite burns, and a cell's byte changes exactly twice in its life.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sfb_rothermel.cuh"

namespace sfb {

constexpr int ST_UNBURNED = 0, ST_BURNING = 1, ST_BURNED = 2, ST_LINE_BIT = 4;
constexpr int DIR_NONE = 8;  // work item of a control-line cell that is not a candidate

struct __align__(16) EnvMeta {
    int32_t t;          // 1-based index of the update() call being executed
    int32_t running;    // GameStatus (1 RUNNING, 0 QUIT)
    int32_t time_quit;  // this step hits the max_time rule (fire.py:641-643)
    int32_t any_live;   // set by k_sweep: a sprite survives pruning (fire.py:637)
    int32_t any_cand;   // set by k_sweep: at least one (src, dst) pair exists (fire.py:651)
    int32_t pad;
    double elapsed;     // elapsed_time (fire.py:717)
};

struct StaticRec {  // 32 B = one DRAM sector
    float4 fuel;    // w_0, delta, M_x, sigma
    float4 env;     // U, U_dir, slope_mag, slope_dir
};

struct DevParams {
    int32_t H, W, E, pitch;  // pitch in cells, multiple of 16
    int32_t max_dur, diagonal, attenuate, shared_static, keep_ros, has_max_time;
    int32_t rows_per_chunk, strips, chunks;
    int64_t n_units;  // E * chunks * strips (one warp each)
    int64_t plane;    // H * pitch
    int64_t qcap;
    double ps, dt, max_time;
    SfbParticle part;
    void* state;
    double* burn;
    double* ros;
    const StaticRec* stat;
    EnvMeta* meta;              // [2][E], double-buffered by step parity
    unsigned long long* queue;  // [qcap] work items
    unsigned long long* qcount; // [2]
    int32_t* overflow;          // [2]
    // slab mode: rows -1 and H of this slab live in a neighbour slab (peer device memory)
    const void* halo_top;     // row (slab_y0 - 1) of the slab above, or nullptr
    const void* halo_bottom;  // row (slab_y0 + H) of the slab below, or nullptr
    const void* filler;       // (pitch + 2 * 16) BURNED cells: stands in for rows outside the grid
};

template <typename CellT>
struct Cell;
template <>
struct Cell<uint8_t> {
    static constexpr int CPL = 16;  // cells per lane per 128-bit load
    static constexpr int M = 31;
    static constexpr uint32_t FILL = 0x02020202u;      // BURNED, no sprite
    static constexpr uint32_t CODE_MASK = 0xF8F8F8F8u; // any sprite code in the word
    static constexpr uint32_t LINE_MASK = 0x04040404u; // any control-line cell in the word
};
template <>
struct Cell<uint16_t> {
    static constexpr int CPL = 8;
    static constexpr int M = 8191;
    static constexpr uint32_t FILL = 0x00020002u;
    static constexpr uint32_t CODE_MASK = 0xFFF8FFF8u;
    static constexpr uint32_t LINE_MASK = 0x00040004u;
};

__device__ __forceinline__ int to_internal(int burn_status) { return burn_status >= 3 ? burn_status + 1 : burn_status; }
__device__ __forceinline__ int to_burn_status(int internal) { return internal >= 4 ? internal - 1 : internal; }
__device__ __forceinline__ bool ignitable(int s) { return s == ST_UNBURNED || (s & ST_LINE_BIT); }

// RoSAttenuation (simfire/enums.py:83-85) by internal status 4, 5, 6
__device__ __forceinline__ double line_attenuation(int s) { return s == 4 ? 980.0 : (s == 5 ? 490.0 : 245.0); }

// duration of a sprite with code `code` as seen by the update() call whose (t-1) mod M is tm1
template <typename CellT>
__device__ __forceinline__ int sprite_age(int code, int tm1) {
    int a = tm1 - (code - 1);
    return a < 0 ? a + Cell<CellT>::M : a;
}

__device__ __forceinline__ unsigned long long make_item(long long idx, int dir, int s) {
    return (unsigned long long)idx | ((unsigned long long)dir << 48) | ((unsigned long long)s << 52);
}

// ---------------------------------------------------------------------------------------
// Work item: the part of the step that touches float data.  Shared by the queue path and
// the dense fallback.
// ---------------------------------------------------------------------------------------
template <typename CellT>
__device__ __forceinline__ void process_item(const DevParams& p, const EnvMeta& m, int env, long long idx,
                                             int dir, int s) {
    double ros;
    if (dir != DIR_NONE) {
        const long long cell = idx - (long long)env * p.plane;
        const StaticRec* rp = p.stat + (p.shared_static ? cell : idx);
        const float4 f = __ldg(&rp->fuel);
        const float4 e = __ldg(&rp->env);
        const float rec[8] = {f.x, f.y, f.z, f.w, e.x, e.y, e.z, e.w};
        ros = sfb_rate_of_spread_pair(dir, rec, p.part) * p.dt;  // fire.py:696
        if (s & ST_LINE_BIT) ros = p.attenuate ? ros - line_attenuation(s) : 0.0;  // fire.py:271-282
    } else {
        // control line that no fire touches: attenuated only if the step got past the
        // "no new locations" early return (fire.py:651-652)
        if (!m.any_cand) return;
        ros = 0.0 - line_attenuation(s);
    }
    if (p.keep_ros) p.ros[idx] = ros;
    double b = p.burn[idx];
    if (ros != 0.0) {  // burn + 0 == burn: skip the store
        b += ros;      // fire.py:710
        p.burn[idx] = b;
    }
    if (dir != DIR_NONE && b > p.ps) {  // fire.py:568 (strict); tested for every candidate
        const int code = 1 + (m.t % Cell<CellT>::M);
        reinterpret_cast<CellT*>(p.state)[idx] = (CellT)(ST_BURNING | (code << 3));  // fire.py:571-587
    }
}

// ---------------------------------------------------------------------------------------
// k_sweep: one warp per (env, chunk of rows, strip of 32 x CPL columns).  Each lane streams
// its CPL-cell segment of every row with one 128-bit load, four rows in flight, keeping a
// three-row window in registers.  A row is examined cell by cell only if the window holds
// a sprite code (or, with attenuation, a control line): the window is then staged in
// shared memory so that the eight neighbours of every cell are plain byte reads.
//
// The streaming part is straight-line code: rows outside the grid are read from a row of
// BURNED filler cells (p.filler) instead of being branched around, and the cells left and
// right of the strip come from one predicated byte load by lanes 0 and 31.
// ---------------------------------------------------------------------------------------
constexpr int SWEEP_WARPS = 4;
constexpr int WQ_CAP = 96;  // >= 64: a flush is forced whenever fewer than 32 slots are free

struct RowRegs {
    uint4 v;     // this lane's CPL cells
    uint32_t h;  // lane 0: cell left of the strip, lane 31: cell right of it
    uint32_t b;  // ballot: lanes whose segment needs a look
};

template <typename CellT>
__global__ void __launch_bounds__(SWEEP_WARPS * 32) k_sweep(const DevParams p, const int par) {
    using C = Cell<CellT>;
    constexpr int CPL = C::CPL;
    constexpr int WR = 32 * CPL;        // cells per warp row
    constexpr int RS = WR + 2 * CPL;    // staged row: CPL pad | WR cells | CPL pad (16-B aligned)
    constexpr int SEG_PER_GROUP = 32 / CPL;
    constexpr uint32_t CELL_ALL = sizeof(CellT) == 1 ? 0xFFu : 0xFFFFu;
    __shared__ __align__(16) CellT sm_all[SWEEP_WARPS][3][RS];
    __shared__ unsigned long long wq_all[SWEEP_WARPS][WQ_CAP];  // per-warp staging of work items

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long unit = (long long)blockIdx.x * SWEEP_WARPS + warp;
    if (unit >= p.n_units) return;
    const int strip = (int)(unit % p.strips);
    const long long u2 = unit / p.strips;
    const int chunk = (int)(u2 % p.chunks);
    const int env = (int)(u2 / p.chunks);
    EnvMeta* const mp = p.meta + (long long)par * p.E + env;
    if (!mp->running) return;
    const bool spread = !mp->time_quit;
    const int tm1 = (mp->t - 1) % C::M;
    const int max_dur = p.max_dur;
    const int H = p.H, pitch = p.pitch;
    const bool diagonal = p.diagonal != 0, attenuate = p.attenuate != 0;

    CellT* const state = reinterpret_cast<CellT*>(p.state);
    const long long env_off = (long long)env * p.plane;
    const CellT* const envbase = state + env_off;
    const CellT* const filler = reinterpret_cast<const CellT*>(p.filler) + CPL;  // valid for [-CPL, pitch + CPL)
    const int x0 = strip * WR;
    const int xl = x0 + lane * CPL;
    const int y_begin = chunk * p.rows_per_chunk;
    const int y_end = min(y_begin + p.rows_per_chunk, H);
    const uint32_t look_mask = attenuate ? (C::CODE_MASK | C::LINE_MASK) : C::CODE_MASK;
    const uint32_t hmask = look_mask & CELL_ALL;
    const bool in_x = xl < pitch;
    // lanes 0 / 31 fetch the cell just outside the strip (if there is one)
    const int hoff = lane == 0 ? x0 - 1 : x0 + WR;
    const bool hpred = (lane == 0 && strip > 0) || (lane == 31 && x0 + WR < pitch);
    CellT(*sm)[RS] = sm_all[warp];
    unsigned long long* const wq = wq_all[warp];
    int wcount = 0;  // warp-uniform

    // rows outside [0, H): the neighbour slab's edge row in slab mode, BURNED filler otherwise
    auto edge_row = [&](int y) -> const CellT* {
        if (y == -1 && p.halo_top) return reinterpret_cast<const CellT*>(p.halo_top) + env_off;
        if (y == H && p.halo_bottom) return reinterpret_cast<const CellT*>(p.halo_bottom) + env_off;
        return filler;
    };
    auto issue = [&](const CellT* rowp, RowRegs& r) {
        r.v = make_uint4(C::FILL, C::FILL, C::FILL, C::FILL);
        r.h = ST_BURNED;
        if (in_x) r.v = *reinterpret_cast<const uint4*>(rowp + xl);
        if (hpred) r.h = rowp[hoff];
    };
    auto finish = [&](RowRegs& r) {
        const uint32_t a = ((r.v.x | r.v.y | r.v.z | r.v.w) & look_mask) | (r.h & hmask);
        r.b = __ballot_sync(0xffffffffu, a != 0);
    };

    // one global atomic per flush instead of one per push: the queue tail is a single
    // address and L2 serialises atomics on it
    auto flush = [&]() {
        if (wcount == 0) return;
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.qcount + par, (unsigned long long)wcount);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < wcount; i += 32) {
            const unsigned long long slot = base + i;
            if (slot < (unsigned long long)p.qcap) p.queue[slot] = wq[i];
            else p.overflow[par] = 1;
        }
        wcount = 0;
        __syncwarp();
    };

    int f_live = 0, f_cand = 0;

    auto process_row = [&](int y, const RowRegs& rp, const RowRegs& rc, const RowRegs& rn) {
        const uint32_t act = rp.b | rc.b | rn.b;
        if (act == 0 || y >= y_end) return;  // warp-uniform
        __syncwarp();
        *reinterpret_cast<uint4*>(&sm[0][CPL + lane * CPL]) = rp.v;
        *reinterpret_cast<uint4*>(&sm[1][CPL + lane * CPL]) = rc.v;
        *reinterpret_cast<uint4*>(&sm[2][CPL + lane * CPL]) = rn.v;
        if (lane == 0) { sm[0][CPL - 1] = (CellT)rp.h; sm[1][CPL - 1] = (CellT)rc.h; sm[2][CPL - 1] = (CellT)rn.h; }
        if (lane == 31) { sm[0][CPL + WR] = (CellT)rp.h; sm[1][CPL + WR] = (CellT)rc.h; sm[2][CPL + WR] = (CellT)rn.h; }
        __syncwarp();
        // segments whose own or adjacent segment needs a look -> 32-cell groups to visit
        const uint32_t near = act | (act << 1) | (act >> 1);
        uint32_t groups = 0;
#pragma unroll
        for (int g = 0; g < WR / 32; ++g)
            if (near & (((1u << SEG_PER_GROUP) - 1u) << (g * SEG_PER_GROUP))) groups |= 1u << g;
        while (groups) {  // warp-uniform
            const int g = __ffs(groups) - 1;
            groups &= groups - 1;
            const int xi = CPL + g * 32 + lane;  // index in the staged row
            const int x = x0 + g * 32 + lane;
            const int c = sm[1][xi];
            int s = c & 7;
            const int code = c >> 3;
            const long long idx = env_off + (long long)y * pitch + x;
            if (code) {
                if (sprite_age<CellT>(code, tm1) >= max_dur) {  // fire.py:116-161
                    state[idx] = (CellT)ST_BURNED;
                    s = ST_BURNED;
                } else {
                    f_live = 1;
                }
            }
            bool push = false;
            int dir = DIR_NONE;
            if (spread && ignitable(s)) {
                // the pair written last is the one whose source has the largest
                // (ignition step, y, x): smallest age, then the order below
                int best = max_dur;
                auto look = [&](int row, int dx, int k) {
                    const int nc = (int)sm[row][xi + dx] >> 3;
                    if (nc) {
                        const int a = sprite_age<CellT>(nc, tm1);
                        if (a < best) { best = a; dir = k; }
                    }
                };
                if (diagonal) look(2, +1, 5);
                look(2, 0, 6);
                if (diagonal) look(2, -1, 7);
                look(1, +1, 4);
                look(1, -1, 0);
                if (diagonal) look(0, +1, 3);
                look(0, 0, 2);
                if (diagonal) look(0, -1, 1);
                if (dir != DIR_NONE) {
                    f_cand = 1;
                    push = true;
                } else if ((s & ST_LINE_BIT) && attenuate) {
                    push = true;
                }
            }
            const uint32_t pm = __ballot_sync(0xffffffffu, push);
            if (pm) {
                if (push) wq[wcount + __popc(pm & ((1u << lane) - 1))] = make_item(idx, dir, s);
                wcount += __popc(pm);
                if (wcount > WQ_CAP - 32) flush();
            }
        }
    };

    // rolling three-row window; the four loads of a batch are issued before any is used
    RowRegs r0, r1, r2, r3, r4, r5;
    const CellT* rowp = envbase + (long long)y_begin * pitch;  // row y of the loop below
    issue(y_begin > 0 ? rowp - pitch : edge_row(-1), r0);
    issue(rowp, r1);
    finish(r0);
    finish(r1);
    for (int y = y_begin; y < y_end; y += 4) {
        const CellT* q1 = rowp + pitch;
        const CellT* q2 = q1 + pitch;
        const CellT* q3 = q2 + pitch;
        const CellT* q4 = q3 + pitch;
        if (y + 4 >= H) {  // warp-uniform, last batch of the grid only
            if (y + 1 >= H) q1 = edge_row(y + 1);
            if (y + 2 >= H) q2 = edge_row(y + 2);
            if (y + 3 >= H) q3 = edge_row(y + 3);
            q4 = edge_row(y + 4);
        }
        issue(q1, r2);
        issue(q2, r3);
        issue(q3, r4);
        issue(q4, r5);
        finish(r2);
        finish(r3);
        finish(r4);
        finish(r5);
        process_row(y, r0, r1, r2);
        process_row(y + 1, r1, r2, r3);
        process_row(y + 2, r2, r3, r4);
        process_row(y + 3, r3, r4, r5);
        r0 = r4;
        r1 = r5;
        rowp = q4;
    }
    flush();
    f_live = __any_sync(0xffffffffu, f_live);
    f_cand = __any_sync(0xffffffffu, f_cand);
    if (lane == 0) {
        if (f_live) mp->any_live = 1;
        if (f_cand) mp->any_cand = 1;
    }
}

// ---------------------------------------------------------------------------------------
// k_eval: persistent grid-stride over the work queue (or, if the queue overflowed, over
// every cell).  Threads 0..E-1 also write the next step's EnvMeta.
// ---------------------------------------------------------------------------------------
template <typename CellT>
__device__ void dense_cell(const DevParams& p, const int par, long long idx) {
    using C = Cell<CellT>;
    const int env = (int)(idx / p.plane);
    const long long cell = idx - (long long)env * p.plane;
    const int y = (int)(cell / p.pitch), x = (int)(cell - (long long)y * p.pitch);
    if (x >= p.W) return;
    const EnvMeta m = p.meta[(long long)par * p.E + env];
    if (!m.running || m.time_quit) return;
    const CellT* st = reinterpret_cast<const CellT*>(p.state);
    const int s = st[idx] & 7;
    if (!ignitable(s)) return;
    const int tm1 = (m.t - 1) % C::M;
    int best = p.max_dur, dir = DIR_NONE;
    auto look = [&](int dy, int dx, int k) {
        const int yy = y + dy, xx = x + dx;
        if (xx < 0 || xx >= p.W) return;
        const CellT* rowp;
        if (yy >= 0 && yy < p.H) rowp = st + (long long)env * p.plane + (long long)yy * p.pitch;
        else if (yy == -1 && p.halo_top) rowp = reinterpret_cast<const CellT*>(p.halo_top) + (long long)env * p.plane;
        else if (yy == p.H && p.halo_bottom) rowp = reinterpret_cast<const CellT*>(p.halo_bottom) + (long long)env * p.plane;
        else return;
        const int nc = (int)rowp[xx] >> 3;
        if (nc) {
            const int a = sprite_age<CellT>(nc, tm1);  // a cell ignited by this very pass has age M-1 >= max_dur
            if (a < best) { best = a; dir = k; }
        }
    };
    if (p.diagonal) look(+1, +1, 5);
    look(+1, 0, 6);
    if (p.diagonal) look(+1, -1, 7);
    look(0, +1, 4);
    look(0, -1, 0);
    if (p.diagonal) look(-1, +1, 3);
    look(-1, 0, 2);
    if (p.diagonal) look(-1, -1, 1);
    if (dir == DIR_NONE && !((s & ST_LINE_BIT) && p.attenuate)) return;
    process_item<CellT>(p, m, env, idx, dir, s);
}

template <typename CellT>
__global__ void __launch_bounds__(256) k_eval(const DevParams p, const int par) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;

    if (p.overflow[par]) {
        const long long total = (long long)p.E * p.plane;
        for (long long i = gid; i < total; i += gstride) dense_cell<CellT>(p, par, i);
    } else {
        const long long n = (long long)p.qcount[par];
        for (long long i = gid; i < n; i += gstride) {
            const unsigned long long it = p.queue[i];
            const long long idx = (long long)(it & 0xFFFFFFFFFFFFull);
            const int dir = (int)((it >> 48) & 0xF), s = (int)((it >> 52) & 7);
            const int env = (int)(idx / p.plane);
            const EnvMeta m = p.meta[(long long)par * p.E + env];
            process_item<CellT>(p, m, env, idx, dir, s);
        }
    }

    // per-env clock for the next step (reads only what k_sweep finalised)
    for (long long env = gid; env < p.E; env += gstride) {
        const EnvMeta cur = p.meta[(long long)par * p.E + env];
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
        p.meta[(long long)(par ^ 1) * p.E + env] = nxt;
    }
    if (gid == 0) {
        p.qcount[par ^ 1] = 0;
        p.overflow[par ^ 1] = 0;
    }
}

}  // namespace sfb
