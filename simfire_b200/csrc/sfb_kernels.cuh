// Device side of the B200 fire-spread stepper: data layout + the three hot-path kernels.
//
// One timestep of RothermelFireManager.update (simfire/game/managers/fire.py:616-719) is a
// front end followed by two front-proportional kernels:
//
//   front end lists the warp-rows (one row of one 512-byte strip of one env) whose 3-row window
//             holds a Fire sprite (or, with attenuation, a control line) as 8-byte row tasks:
//     k_row_list  (default from 1024 units up) compacts per-(env, row, strip) activity flags that
//                 the other kernels keep up to date; no cell state is read at all;
//     k_sweep     streaming pass over the packed per-cell state (1 B/cell, TMA-fed), over every
//                 unit (dense: HBM-bound, at the copy roofline) or over the flagged chunks of
//                 rows only (k_units compacts those flags first).  Writes nothing else.
//   k_rows    one warp per row task: prune expired sprites (fire.py:116-161), find every
//             ignitable cell that has a burning neighbour and the neighbour whose pair the
//             reference writes last (fire.py:163-234, :704-705) with a warp-shuffle min, push
//             (cell, direction) work items to a queue, raise the per-env flags the reference's
//             early returns depend on (fire.py:637, :651).  Issue-bound, front-proportional.
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
#ifdef SFB_EMU
// test-only build of these sources with g++ (tests/emu): cuda_emu.h stands in for the CUDA headers
#include "cuda_emu.h"
#else
#include <cuda.h>
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "sfb_rothermel.cuh"

namespace sfb {

constexpr int ST_UNBURNED = 0, ST_BURNING = 1, ST_BURNED = 2, ST_LINE_BIT = 4;
constexpr int DIR_NONE = 8;    // work item of a control-line cell that is not a candidate
constexpr int DIR_PRUNED = 9;  // not work: a sprite burnt out here (only pushed for the change log)

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

// Slab mode: what the slabs of one grid tell each other every step, written straight into the
// peers' memory (NVLink peer stores when the slabs are on different GPUs) and polled by tiny
// single-warp kernels, so a run of n steps is enqueued once and needs no host or NCCL round trip.
constexpr int SLAB_MAX_WORLD = 8, SLAB_MAX_ENVS = 16;
struct SlabMailbox {
    volatile uint32_t done_step[SLAB_MAX_WORLD];   // [q]: slab q has finished k_eval of step ...
    // double-buffered by step parity: a slab that is one step ahead posts into the other half
    // while a slower one is still reading this one (it cannot get two steps ahead: its next
    // exchange needs the slower slab's post)
    volatile uint32_t flag_step[2][SLAB_MAX_WORLD];   // [g & 1][q]: flags[g & 1][q] belong to step g
    volatile int32_t flags[2][SLAB_MAX_WORLD][SLAB_MAX_ENVS][2];  // any_live, any_cand of slab q
    volatile int32_t error;                        // a wait gave up (peer stalled)
};

constexpr int MAX_ENV_GROUPS = 16;
struct LogRef {
    unsigned long long* buf;    // entries (host-mapped memory)
    unsigned long long* count;  // [0] entries so far, [1] overflow flag
    long long cap;
};

struct DerivedRec {  // 48 B: what k_eval gathers per candidate item
    SfbFuelTerms fuel;  // fuel-only Rothermel terms of the cell (k_derive_static)
    float4 env;         // U, U_dir, slope_mag, slope_dir
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
    double* ros_w;           // where a step writes its rates: `ros`, or (bitboard handles) a plane committed after k_eval
    int32_t* ign;            // update() call that ignited the cell (SFB_KEEP_IGNITION), else nullptr
    const StaticRec* stat;   // raw inputs as uploaded
    const DerivedRec* drv;   // derived from `stat` by k_derive_static before the first step that needs it
    EnvMeta* meta;              // [2][meta_stride], double-buffered by step parity
    int32_t meta_stride;        // envs of the whole handle (a kernel may see a group of them: E <= meta_stride)
    int64_t idx_base;           // cell index of this group's first cell within the handle (change-log entries)
    unsigned long long* queue;  // [qcap] work items
    unsigned long long* qcount; // [2]
    int32_t* overflow;          // [2]
    unsigned long long* unit_next;  // [2] next sweep unit to hand out (dynamic scheduling)
    unsigned long long* rows;       // [rows_cap] row tasks written by k_sweep, consumed by k_rows
    unsigned long long* rows_count; // [2]
    unsigned long long* rows_next;  // [2] next row task to hand out
    int64_t rows_cap;
    // unit skipping: unit_act[u] != 0 <=> the sweep has to read unit u = (env, chunk, strip).  Invariant:
    // a unit whose rows / columns, or the one-cell frame around them, hold a sprite code (or, with
    // attenuation, a control line) is flagged.  Flags are raised wherever a cell ignites or a control
    // line is drawn (process_item, k_reset_meta, k_apply_points; sfb_set_fire_map flags whole envs) and
    // lowered by the sweep itself when a flagged unit yields no row task.  k_units compacts the
    // flagged units of running envs into `units`; the sweep then only draws from that list.
    // Row units (unit_rows != 0): a unit is one row of one strip (rows_per_chunk = 1, chunks = H).  Then the
    // flags ARE the row-task list: k_row_list compacts them straight into `rows`, no state is swept at
    // all, and k_rows lowers the flag of a row whose 3-row window holds nothing to look at.
    int32_t unit_rows, unit_pad_;
    int64_t unit_stride;              // flags per env: chunks * strips (row units: rounded up to 4, pad flags unused)
    uint8_t* unit_act;                // [E * chunks * strips], nullptr = dense sweep over every unit
    uint32_t* units;                  // [n_units] this step's active units
    unsigned long long* units_count;  // [2]
    // slab mode: rows -1 and H of this slab live in a neighbour slab (peer device memory)
    const void* halo_top;     // row (slab_y0 - 1) of the slab above, or nullptr
    const void* halo_bottom;  // row (slab_y0 + H) of the slab below, or nullptr
    int64_t halo_top_plane, halo_bottom_plane;  // per-env stride (cells) of those neighbour planes
    const void* filler;       // (pitch + 2 * 16) BURNED cells: stands in for rows outside the grid
    // slab mode without host round trips: per-step agreement through peer-visible mailboxes
    struct SlabMailbox* mailbox;       // this slab's mailbox (peers write into it)
    struct SlabMailbox* peer_box[8];   // every slab's mailbox as seen from this device (own included)
    int32_t slab_rank, slab_world;
    // change log (SFB_TRACK_CHANGES): one log per env group so that each stays ordered by env and
    // the host can patch the log of a group that is done while the other groups still compute.
    // A view's k_eval appends to chg (its own group's log); the setup kernels, which see the
    // whole handle, route by env through logs[] / log_e0[].
    int32_t track;
    unsigned long long* chg;        // [chg_cap] idx | BurnStatus << 48
    unsigned long long* chg_count;  // entries appended since the host last drained the log
    int32_t* chg_overflow;
    int64_t chg_cap;
    int32_t n_logs;
    int32_t log_e0[MAX_ENV_GROUPS + 1];  // log g holds envs [log_e0[g], log_e0[g + 1])
    LogRef logs[MAX_ENV_GROUPS];
    // list-driven step (sfb_lists.cuh): one list of the cells that can change in a step
    unsigned long long* wl[2];       // watch-list buffers; k_front reads wl[lpar] and writes wl[lpar ^ 1]
    unsigned long long* wl_count;    // [2] entries of wl[k]
    int64_t wl_cap;
    uint32_t* listed;                // one bit per cell: the cell has an entry (nullptr: not a list handle)
    int32_t* broken;                 // sticky: an append found the list full -> dense form of the same step
    int32_t* dense_now;              // ... latched between launches: the form the next k_front takes
    unsigned int* ticket;            // blocks of k_front that are done (the last one closes the step)
    const double* rtab;              // [static cells][8] rate of spread per direction (ft/min), or nullptr
    unsigned long long* ros_items;   // SFB_KEEP_ROS: (cell index, float64 bits) of this step's candidates
    unsigned long long* ros_count;
    int64_t ros_cap;
    // bitboard front end (sfb_bits.cuh): bit planes next to the state bytes, one word per (env, tile column, row)
    uint32_t* bits;                  // [E][2 + ring][tiles_x][H] (nullptr: not a bitboard handle)
    int32_t ring, tiles_x, tiles_y, bits_pad_;  // tiles_x = ceil(W / 30): a word owns 30 cells
    int64_t bits_plane, bits_env;    // words per plane (tiles_x * H) and per env
    uint8_t* tile_act;               // [2][E][tile_stride] activity flag per tile of 32 rows x 30 columns, by step parity
    int64_t tile_stride, tile_buf;   // flags per env (a multiple of 16) and per buffer (all envs of the handle)
    int32_t bits_par, bits_pad2_;    // setup kernels: the buffer the next step reads
    unsigned long long* tile_stats;  // [FRONT_N_STATS] accumulated by k_tiles while kernel timing is on, else nullptr
    unsigned long long* late;        // cells that re-ignite while still a source for this step (rewritten by the last block)
    unsigned int* late_count;
    int64_t late_cap;
    unsigned long long* front_stats; // [FRONT_N_STATS] accumulated by k_front, read and reset by sfb_get_front_stats
};

// kernels that may be launched with programmatic stream serialization (SFB_LAUNCH_DEP): nothing the preceding
// kernel of the stream wrote is read before this returns; a no-op for ordinary launches
__device__ __forceinline__ void grid_dep_wait() {
#ifndef SFB_EMU
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// ... and the successor may be scheduled from here on (it still waits for this grid to finish)
__device__ __forceinline__ void grid_dep_launch() {
#ifndef SFB_EMU
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ const LogRef& log_of_env(const DevParams& p, int env) {
    int g = 0;
    while (g + 1 < p.n_logs && env >= p.log_e0[g + 1]) ++g;
    return p.logs[g];
}
__device__ __forceinline__ void log_put(const LogRef& L, unsigned long long slot, unsigned long long entry) {
    if (slot < (unsigned long long)L.cap) L.buf[slot] = entry;
    else *reinterpret_cast<int32_t*>(L.count + 1) = 1;
}

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

// raise the activity flag of every unit that holds cell (y, x) of `env` in its rows / columns or in
// the one-cell frame around them (at most 2 x 2 units)
template <typename CellT>
__device__ __forceinline__ void mark_units_around(const DevParams& p, int env, int y, int x) {
    constexpr int WR = 32 * Cell<CellT>::CPL;
    const int R = p.rows_per_chunk;
    const int y0 = y > 0 ? y - 1 : 0, y1 = y + 1 < p.H ? y + 1 : p.H - 1;
    const int x0 = x > 0 ? x - 1 : 0, x1 = x + 1 < p.W ? x + 1 : p.W - 1;
    const int c0 = R == 1 ? y0 : y0 / R, c1 = R == 1 ? y1 : (y1 >= (c0 + 1) * R ? c0 + 1 : c0);  // y1 - y0 <= 2
    const int s0 = x0 / WR, s1 = x1 / WR;
    uint8_t* f = p.unit_act + (long long)env * p.unit_stride + (long long)c0 * p.strips;
    for (int c = c0; c <= c1; ++c, f += p.strips) {  // at most three rows of flags (row units), else two
        f[s0] = 1;
        if (s1 != s0) f[s1] = 1;
    }
}
// the same from a cell index relative to the view's first cell
template <typename CellT>
__device__ __forceinline__ void mark_units_of_cell(const DevParams& p, int env, long long idx) {
    const long long cell = idx - (long long)env * p.plane;
    int y, x;
    if (p.plane <= 0x7fffffffll) {  // warp-uniform; 32-bit division
        const uint32_t c = (uint32_t)cell;
        y = (int)(c / (uint32_t)p.pitch);
        x = (int)(c - (uint32_t)y * (uint32_t)p.pitch);
    } else {
        y = (int)(cell / p.pitch);
        x = (int)(cell - (long long)y * p.pitch);
    }
    mark_units_around<CellT>(p, env, y, x);
}

__device__ __forceinline__ unsigned long long make_item(long long idx, int dir, int s) {
    return (unsigned long long)idx | ((unsigned long long)dir << 48) | ((unsigned long long)s << 52);
}

// ---------------------------------------------------------------------------------------
// Work item: the part of the step that touches float data.  Shared by the queue path and
// the dense fallback.
// ---------------------------------------------------------------------------------------
// Change log (SFB_TRACK_CHANGES): every status change of a cell is appended as
// idx | BurnStatus << 48 so that a host mirror of fire_map can be patched instead of
// re-downloaded (sfb_sync_fire_maps).  BurnStatus 7 = "env idx was reset".  Warp-wide call.
constexpr int LOG_ENV_RESET = 7;
// bit 56 of a log entry: written by a between-step kernel (reset, mitigation) rather than by a step.
// When one step follows them the host applies these few entries first, in order, and then patches
// the step's entries -- no cell appears twice among those -- in parallel without sorting them.
constexpr unsigned long long LOG_SETUP_BIT = 1ull << 56;
__device__ __forceinline__ void log_append(const DevParams& p, bool have, long long idx, int burn_status) {
    const uint32_t m = __ballot_sync(0xffffffffu, have);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    unsigned long long base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(p.chg_count, (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (have) {
        const unsigned long long slot = base + __popc(m & ((1u << lane) - 1));
        if (slot < (unsigned long long)p.chg_cap) p.chg[slot] = (unsigned long long)idx | ((unsigned long long)burn_status << 48);
        else *p.chg_overflow = 1;
    }
}

// handles without a rate table (it did not fit): the rate from the derived record, float64 pow / cos per item.
// Not inlined: the callers' common path (one table read) should not carry its registers.
#ifdef SFB_EMU
#define SFB_NOINLINE __attribute__((noinline))
#else
#define SFB_NOINLINE __noinline__
#endif
static __device__ SFB_NOINLINE double rate_from_record(const DevParams& p, long long sc, int dir) {
    const float4* rp = reinterpret_cast<const float4*>(p.drv + sc);
    const float4 t0 = __ldg(rp), t1 = __ldg(rp + 1), e = __ldg(rp + 2);
    const SfbFuelTerms t = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
    return sfb_spread_from_terms(dir, t, e.x, e.y, e.z, e.w);
}

// returns true if the cell ignited
template <typename CellT>
__device__ __forceinline__ bool process_item(const DevParams& p, const EnvMeta& m, int env, long long idx,
                                             int dir, int s) {
    double ros;
    if (dir != DIR_NONE) {
        const long long cell = idx - (long long)env * p.plane;
        const long long sc = p.shared_static ? cell : idx;
        if (p.rtab) {
            // every input of rothermel.py:4-136 is static per (cell, direction): k_derive_static evaluated
            // the eight float64 rates of the cell once, through the same device function as below
            ros = __ldg(p.rtab + sc * 8 + dir) * p.dt;  // fire.py:696
        } else {
            ros = rate_from_record(p, sc, dir) * p.dt;  // rothermel.py:4-136, fire.py:696
        }
        if (s & ST_LINE_BIT) ros = p.attenuate ? ros - line_attenuation(s) : 0.0;  // fire.py:271-282
    } else {
        // control line that no fire touches: attenuated only if the step got past the
        // "no new locations" early return (fire.py:651-652)
        if (!m.any_cand) return false;
        ros = 0.0 - line_attenuation(s);
    }
    if (p.keep_ros) p.ros_w[idx] = ros;
    double b = p.burn[idx];
    if (ros != 0.0) {  // burn + 0 == burn: skip the store
        b += ros;      // fire.py:710
        p.burn[idx] = b;
    }
    if (dir != DIR_NONE && b > p.ps) {  // fire.py:568 (strict); tested for every candidate
        const int code = 1 + (m.t % Cell<CellT>::M);
        reinterpret_cast<CellT*>(p.state)[idx] = (CellT)(ST_BURNING | (code << 3));  // fire.py:571-587
        if (p.ign) p.ign[idx] = m.t;
        if (p.unit_act) mark_units_of_cell<CellT>(p, env, idx);
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// The sweep is split in two kernels so that the expensive part is load-balanced:
//
//   k_sweep_*  pure streaming: one persistent warp per (env, chunk of rows, strip of 512 B of
//              columns) unit; decides per warp-row (one AND per 32-bit word) whether its 3-row
//              window holds a sprite code (or, with attenuation, a control line) and, if so,
//              appends an 8-byte row task to a list.  Nothing else: no state writes.
//   k_rows     one warp per row task: re-reads the three rows (L2 / L1 hits: they were just
//              streamed), stages them in shared memory as CPL pad | 32*CPL cells | CPL pad so that
//              the eight neighbours of every cell are plain byte reads, and examines the row cell
//              by cell (RowWorker::detail_row): prune, candidate search with a warp-shuffle min,
//              work items for k_eval.  Row tasks cost about the same, so a fire front that sits
//              in a few units no longer serialises inside the few warps that own them.
//
// Two streaming front ends:
//   k_sweep_tma  TMA (cp.async.bulk.tensor) boxes of 8 rows x 544 B land in a per-warp
//                shared-memory ring, completion on mbarriers, all stages in flight;
//                out-of-grid cells are zero-filled by the TMA unit.  Default.
//   k_sweep_ldg  128-bit global loads, four rows in flight; rows outside the grid are read from
//                a row of BURNED filler cells (or from the neighbour slab in slab mode).
// ---------------------------------------------------------------------------------------
#ifndef SFB_SWEEP_WARPS
#define SFB_SWEEP_WARPS 4
#endif
#ifndef SFB_LDG_MIN_BLOCKS
#define SFB_LDG_MIN_BLOCKS 8
#endif
constexpr int SWEEP_WARPS = SFB_SWEEP_WARPS;
constexpr int WQ_CAP = 96;  // >= 64: a flush is forced whenever fewer than 32 slots are free

// row task: y | strip << 20 | env << 28
__device__ __forceinline__ unsigned long long make_row_task(int env, int y, int strip) {
    return (unsigned long long)(unsigned)y | ((unsigned long long)(unsigned)strip << 20) | ((unsigned long long)(unsigned)env << 28);
}

// per-warp staging of row tasks in shared memory, one global atomic per flush
struct RowTaskList {
    const DevParams& p;
    int par, lane;
    unsigned long long* buf;  // [WQ_CAP] shared
    int count = 0;            // warp-uniform
    __device__ __forceinline__ RowTaskList(const DevParams& p_, int par_, int lane_, unsigned long long* buf_)
        : p(p_), par(par_), lane(lane_), buf(buf_) {}
    // every lane calls; lanes with `have` contribute one task; returns the ballot of `have`
    __device__ __forceinline__ uint32_t push(bool have, unsigned long long task) {
        const uint32_t m = __ballot_sync(0xffffffffu, have);
        if (!m) return 0;
        if (have) buf[count + __popc(m & ((1u << lane) - 1))] = task;
        count += __popc(m);
        if (count > WQ_CAP - 32) flush();
        return m;
    }
    __device__ __forceinline__ void flush() {
        if (count == 0) return;
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.rows_count + par, (unsigned long long)count);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < count; i += 32)
            if (base + i < (unsigned long long)p.rows_cap) p.rows[base + i] = buf[i];
        count = 0;
        __syncwarp();
    }
};

template <typename CellT>
struct RowWorker {
    using C = Cell<CellT>;
    static constexpr int CPL = C::CPL;
    static constexpr int WR = 32 * CPL;      // cells per warp row
    static constexpr int RS = WR + 2 * CPL;  // staged row, in cells (544 bytes)
    static constexpr int SEG_PER_GROUP = 32 / CPL;
    static constexpr uint32_t CELL_ALL = sizeof(CellT) == 1 ? 0xFFu : 0xFFFFu;
    static constexpr int GW = 30, NG = (WR + GW - 1) / GW;  // detail_row visits groups of 30 owned cells

    // seg_lut[L]: the groups whose columns [30g - 1, 30g + 30] overlap lane L's segment (the pad
    // cells count as part of segments 0 / 31).  Filled once per (persistent) warp.
    static __device__ __forceinline__ void build_seg_lut(uint32_t* lut, int lane) {
        uint32_t m = 0;
        for (int g = 0; g < NG; ++g) {
            const int lo = (GW * g - 1) < 0 ? 0 : (GW * g - 1) / CPL;
            const int hi = (GW * g + GW) / CPL > 31 ? 31 : (GW * g + GW) / CPL;
            if (lane >= lo && lane <= hi) m |= 1u << g;
        }
        lut[lane] = m;
        __syncwarp();
    }

    const DevParams& p;
    int par, lane, env = -1, x0 = 0, tm1 = 0, max_dur, W;
    bool spread = false, diagonal, attenuate, track, lane_owner;
    long long env_off = 0;
    uint32_t look_mask;
    unsigned long long* wq;
    const uint32_t* seg_lut;
    int* key_lut;    // [32] behind seg_lut
    int wcount = 0;  // warp-uniform
    int f_live = 0, f_cand = 0;
    // experimental variants of k_rows, built and measured separately:
    // -DSFB_ROWS_REDUX (group mask with one REDUX), -DSFB_ROWS_PADVEC (16-byte pad vectors and per-env
    // cached row pointers), -DSFB_ROWS_V2 = both
#if defined(SFB_ROWS_V2) && !defined(SFB_ROWS_REDUX)
#define SFB_ROWS_REDUX
#endif
#if defined(SFB_ROWS_V2) && !defined(SFB_ROWS_PADVEC)
#define SFB_ROWS_PADVEC
#endif
#ifdef SFB_ROWS_REDUX
    uint32_t my_groups = 0;             // seg_lut[lane], kept in a register
#endif
#ifdef SFB_ROWS_PADVEC
    const CellT* envbase = nullptr;     // first cell of the current env
    const CellT* row_above = nullptr;   // what stands in for row -1 / row H of the current env
    const CellT* row_below = nullptr;
#endif

    __device__ __forceinline__ RowWorker(const DevParams& p_, int par_, int lane_, unsigned long long* wq_,
                                         const uint32_t* lut_)
        : p(p_), par(par_), lane(lane_), wq(wq_), seg_lut(lut_),
          key_lut(reinterpret_cast<int*>(const_cast<uint32_t*>(lut_)) + 32) {
        max_dur = p.max_dur;
        W = p.W;
        diagonal = p.diagonal != 0;
        attenuate = p.attenuate != 0;
        track = p.track != 0;
        lane_owner = lane >= 1 && lane <= GW;
        look_mask = attenuate ? (C::CODE_MASK | C::LINE_MASK) : C::CODE_MASK;
    }

    // switch to another env: publish the flags gathered for the previous one, rebuild what
    // depends on the env's clock
    __device__ __forceinline__ void set_env(int env_, const EnvMeta& m) {
        publish_flags();
        env = env_;
        tm1 = (m.t - 1) % C::M;
        spread = !m.time_quit;
        env_off = (long long)env * p.plane;
        build_key_lut();
#ifdef SFB_ROWS_PADVEC
        const CellT* const filler = reinterpret_cast<const CellT*>(p.filler) + CPL;
        envbase = reinterpret_cast<const CellT*>(p.state) + env_off;
        row_above = p.halo_top ? reinterpret_cast<const CellT*>(p.halo_top) + (long long)env * p.halo_top_plane : filler;
        row_below = p.halo_bottom ? reinterpret_cast<const CellT*>(p.halo_bottom) + (long long)env * p.halo_bottom_plane : filler;
#endif
    }

    __device__ __forceinline__ void publish_flags() {
        f_live = __any_sync(0xffffffffu, f_live);
        f_cand = __any_sync(0xffffffffu, f_cand);
        if (env >= 0 && lane == 0) {
            EnvMeta* mp = p.meta + (long long)par * p.meta_stride + env;
            if (f_live) mp->any_live = 1;
            if (f_cand) mp->any_cand = 1;
        }
        f_live = f_cand = 0;
    }

    // does this lane's segment (plus, for lanes 0 / 31, the cell outside the strip) need a look?
    __device__ __forceinline__ bool seg_needs_look(const uint4& v, uint32_t halo_cell) const {
        return ((((v.x | v.y) | (v.z | v.w)) & look_mask) | (halo_cell & look_mask & CELL_ALL)) != 0;
    }

    // one global atomic per flush instead of one per push: the queue tail is a single
    // address and L2 serialises atomics on it
    __device__ __forceinline__ void flush() {
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
    }

    // sort key of a cell as a fire source: duration * 8 if it carries a live sprite, else NO_SRC
    static constexpr int NO_SRC = 1 << 20;
    __device__ __forceinline__ int source_key_of_code(int code) const {
        const int a = sprite_age<CellT>(code, tm1);
        return (code != 0 && a < max_dur) ? a * 8 : NO_SRC;
    }
    // 8-bit cells have 32 sprite codes: one table look-up (filled per unit, it depends on t)
    __device__ __forceinline__ int source_key(int c) const {
        if constexpr (sizeof(CellT) == 1) return key_lut[c >> 3];
        else return source_key_of_code(c >> 3);
    }
    __device__ __forceinline__ void build_key_lut() {
        if constexpr (sizeof(CellT) == 1) {
            __syncwarp();
            key_lut[lane] = source_key_of_code(lane);
            __syncwarp();
        }
    }

    // rp / rc / rn: rows y-1, y, y+1 in shared memory (RS cells each); act: ballot of the
    // lanes whose segment needs a look in any of the three rows.
    //
    // The row is visited in groups of 30 cells: lane l of group g holds column 30g + l - 1 of
    // the three rows, so lanes 1..30 find all eight neighbours in the adjacent lanes (lanes 0
    // and 31 only lend their cells).  The pair the reference writes last is the one whose
    // source has the largest (ignition step, y, x) (fire.py:704-705 + sprite-list order):
    // smallest duration first, then south before north and east before west.  Folding that
    // rank into the low bits of the key turns the selection into a warp-shuffle min.
    __device__ __forceinline__ void detail_row(int y, const CellT* rp, const CellT* rc, const CellT* rn, uint32_t act) {
        CellT* const state = reinterpret_cast<CellT*>(p.state);
        const long long row_idx = env_off + (long long)y * p.pitch;  // cell index of (y, x = 0)
        const int cells = min(W - x0, WR);  // columns of this strip that exist in the grid
#ifdef SFB_ROWS_REDUX
        uint32_t groups = __reduce_or_sync(0xffffffffu, ((act >> lane) & 1u) ? my_groups : 0u);  // one REDUX
#else
        uint32_t groups = 0;
        for (uint32_t a = act; a; a &= a - 1) groups |= seg_lut[__ffs(a) - 1];  // warp-uniform
#endif
        while (groups) {  // warp-uniform
            const int g = __ffs(groups) - 1;
            groups &= groups - 1;
            const int col = min(g * GW + lane - 1, WR);  // staged column -1 .. WR (pads included)
            const int xi = CPL + col;
            const int cp = rp[xi], cc = rc[xi], cn = rn[xi];
            if (!__any_sync(0xffffffffu, ((cp | cc | cn) & (int)(look_mask & CELL_ALL)) != 0)) continue;
            const int kp = source_key(cp), kc = source_key(cc), kn = source_key(cn);
            // keys of the eight neighbours, rank in the low three bits (0 = written last).  Each lane first
            // folds its own column as it looks from the cell to its west (ranks SE 0, E 3, NE 5) and from
            // the cell to its east (SW 2, W 4, NW 7); the neighbours then fetch one value each: two
            // shuffles instead of six.
#ifdef SFB_ROWS_SHFL6
            int best = min(__shfl_down_sync(0xffffffffu, kc, 1) + 3, __shfl_up_sync(0xffffffffu, kc, 1) + 4);
            best = min(best, min(kn + 1, kp + 6));
            if (diagonal) {
                best = min(best, min(__shfl_down_sync(0xffffffffu, kn, 1) + 0, __shfl_up_sync(0xffffffffu, kn, 1) + 2));
                best = min(best, min(__shfl_down_sync(0xffffffffu, kp, 1) + 5, __shfl_up_sync(0xffffffffu, kp, 1) + 7));
            }
#else
            int as_east = kc + 3, as_west = kc + 4;
            if (diagonal) {
                as_east = min(as_east, min(kn + 0, kp + 5));
                as_west = min(as_west, min(kn + 2, kp + 7));
            }
            int best = min(__shfl_down_sync(0xffffffffu, as_east, 1), __shfl_up_sync(0xffffffffu, as_west, 1));
            best = min(best, min(kn + 1, kp + 6));
#endif
            const int x = x0 + col;
            const bool owner = lane_owner && col < cells;
            int s = cc & 7;
            if (owner && (cc >> 3) != 0) {
                if (kc == NO_SRC) {  // duration reached max_fire_duration (fire.py:116-161)
                    state[row_idx + x] = (CellT)ST_BURNED;
                    s = ST_BURNED;
                } else {
                    f_live = 1;
                }
            }
            bool push = false;
            int dir = DIR_NONE;
            if (track && owner && (cc >> 3) != 0 && kc == NO_SRC) {
                push = true;
                dir = DIR_PRUNED;
            }
            if (owner && spread && ignitable(s)) {
                if (best < NO_SRC) {
                    dir = (0x12304765u >> ((best & 7) * 4)) & 0xF;  // rank -> direction of fire.py:211-221
                    f_cand = 1;
                    push = true;
                } else if ((s & ST_LINE_BIT) && attenuate) {
                    push = true;
                }
            }
            const uint32_t pm = __ballot_sync(0xffffffffu, push);
            if (pm) {
                if (push) wq[wcount + __popc(pm & ((1u << lane) - 1))] = make_item(row_idx + x, dir, s);
                wcount += __popc(pm);
                if (wcount > WQ_CAP - 32) flush();
            }
        }
    }

    __device__ __forceinline__ void finish() {
        flush();
        publish_flags();
    }
};


// Warps are persistent: each pulls the next (env, chunk, strip) unit from a device counter,
// so a warp that drew a short unit immediately takes another one.
__device__ __forceinline__ bool next_unit(const DevParams& p, int par, int lane, int& strip, int& chunk, int& env,
                                          long long& unit_id) {
    unsigned long long unit = ~0ull;
    if (lane == 0) {
        const unsigned long long i = atomicAdd(p.unit_next + par, 1ULL);
        if (!p.unit_act) unit = i;                               // dense: every unit in turn
        else if (i < p.units_count[par]) unit = p.units[i];      // only the units k_units listed
    }
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= (unsigned long long)p.n_units) return false;
    unit_id = (long long)unit;
    strip = (int)(unit % p.strips);
    const long long u2 = unit / p.strips;
    chunk = (int)(u2 % p.chunks);
    env = (int)(u2 / p.chunks);
    return true;
}
// a flagged unit that yielded no row task has nothing to look at within one cell of its rows and
// columns: the sweep skips it from the next step on, until a neighbouring ignition flags it again
__device__ __forceinline__ void retire_unit(const DevParams& p, int lane, long long unit_id, uint32_t emitted) {
    if (p.unit_act && !emitted && lane == 0) p.unit_act[unit_id] = 0;
}

// k_units: compacts the flagged units of running envs into this step's unit list
__global__ void __launch_bounds__(256) k_units(const DevParams p, const int par) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long upe = (long long)p.chunks * p.strips;  // units per env
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; base < p.n_units; base += stride) {
        const long long u = base + lane;
        bool act = false;
        if (u < p.n_units && p.unit_act[u]) act = p.meta[(long long)par * p.meta_stride + (int)(u / upe)].running != 0;
        const uint32_t m = __ballot_sync(0xffffffffu, act);
        if (!m) continue;
        unsigned long long slot = 0;
        if (lane == __ffs(m) - 1) slot = atomicAdd(p.units_count + par, (unsigned long long)__popc(m));
        slot = __shfl_sync(0xffffffffu, slot, __ffs(m) - 1);
        if (act) p.units[slot + __popc(m & ((1u << lane) - 1))] = (uint32_t)u;
    }
}

// k_row_list (row units): the flags are per (env, row, strip), so the flagged units ARE this step's row
// tasks.  Each lane takes four flags (one 32-bit load), the warp scans the counts, one atomic per warp.
__global__ void __launch_bounds__(256) k_row_list(const DevParams p, const int par) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = (long long)p.E * p.unit_stride / 4;  // unit_stride is a multiple of 4
    const long long upe = p.unit_stride, used = (long long)p.H * p.strips;
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(p.unit_act);
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n4; base += stride) {
        const long long w = base + lane;
        uint32_t f = w < n4 ? flags[w] : 0u;
        if (!__any_sync(0xffffffffu, f != 0)) continue;
        unsigned long long task[4];
        int cnt = 0;
        if (f) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (!((f >> (8 * b)) & 0xFFu)) continue;
                const long long u = 4 * w + b;
                const int env = (int)(u / upe);
                if (!p.meta[(long long)par * p.meta_stride + env].running) continue;
                const long long r = u - (long long)env * upe;
                if (r >= used) continue;  // pad flag (a map upload sets whole envs, pads included)
                const int y = (int)(r / p.strips), strip = (int)(r - (long long)y * p.strips);
                task[cnt++] = make_row_task(env, y, strip);
            }
        }
        int incl = cnt;  // inclusive scan of the per-lane counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        if (total == 0) continue;
        unsigned long long slot = 0;
        if (lane == 0) slot = atomicAdd(p.rows_count + par, (unsigned long long)total);
        slot = __shfl_sync(0xffffffffu, slot, 0) + (unsigned long long)(incl - cnt);
        for (int i = 0; i < cnt; ++i)
            if (slot + i < (unsigned long long)p.rows_cap) p.rows[slot + i] = task[i];
    }
}

// ---- front end 1: TMA ring --------------------------------------------------------------
#ifndef SFB_TMA_BOX_ROWS
#define SFB_TMA_BOX_ROWS 8
#endif
constexpr int TMA_BOX_ROWS = SFB_TMA_BOX_ROWS;       // rows per TMA box
#ifndef SFB_TMA_STAGES
#define SFB_TMA_STAGES 3
#endif
constexpr int TMA_STAGES = SFB_TMA_STAGES;           // boxes in the per-warp ring
constexpr int TMA_ROW_BYTES = 544;                   // 16 B pad | 512 B | 16 B pad
constexpr int TMA_BOX_BYTES = TMA_BOX_ROWS * TMA_ROW_BYTES;
constexpr int TMA_RING_ROWS = TMA_BOX_ROWS * TMA_STAGES;
constexpr int TMA_WARP_SMEM = TMA_RING_ROWS * TMA_ROW_BYTES + WQ_CAP * 8 + 128;  // ring | row tasks | mbarriers
static_assert(TMA_WARP_SMEM % 128 == 0 && TMA_BOX_BYTES % 128 == 0, "TMA destinations must stay 128-byte aligned");
constexpr int TMA_BLOCK_SMEM = SWEEP_WARPS * TMA_WARP_SMEM + 128;              // + alignment slack

#ifdef SFB_EMU
using emu::smem_u32;
using emu::mbar_init;
using emu::mbar_init_fence;
using emu::mbar_expect_tx;
using emu::mbar_wait;
using emu::tma_load_3d;
#define SFB_DYNAMIC_SMEM(name) unsigned char* const name = emu::st().dyn_smem
#else
#define SFB_DYNAMIC_SMEM(name) extern __shared__ unsigned char name[]
__device__ __forceinline__ uint32_t smem_u32(const void* ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_init_fence() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif

// The tensor map describes the state plane as uint32 [E][H][pitch_bytes / 4]; a box is
// 136 x 8 x 1 elements = 8 rows of 544 bytes starting 16 bytes left of the strip.
template <typename CellT>
__global__ void __launch_bounds__(SWEEP_WARPS * 32)
k_sweep_tma(const __grid_constant__ CUtensorMap tmap, const DevParams p, const int par) {
    using C = Cell<CellT>;
    constexpr int CPL = C::CPL, WR = 32 * CPL, RS = WR + 2 * CPL;
    constexpr int B = TMA_BOX_ROWS, NR = TMA_RING_ROWS;
    constexpr uint32_t CELL_ALL = sizeof(CellT) == 1 ? 0xFFu : 0xFFFFu;
    static_assert(RS * sizeof(CellT) == TMA_ROW_BYTES, "row bytes");
    static_assert(B <= 16, "halo ballot uses lanes 0 .. 2B-1");
    SFB_DYNAMIC_SMEM(smem_raw);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* base = smem_raw + ((128 - (smem_u32(smem_raw) & 127)) & 127) + warp * TMA_WARP_SMEM;
    CellT* const ring = reinterpret_cast<CellT*>(base);
    const uint32_t bar0 = smem_u32(base + NR * TMA_ROW_BYTES + WQ_CAP * 8);
    const uint32_t ring_u32 = smem_u32(ring);
    RowTaskList tasks(p, par, lane, reinterpret_cast<unsigned long long*>(base + NR * TMA_ROW_BYTES));
    const uint32_t look_mask = p.attenuate ? (C::CODE_MASK | C::LINE_MASK) : C::CODE_MASK;

    if (lane == 0) {
        for (int s = 0; s < TMA_STAGES; ++s) mbar_init(bar0 + 8 * s, 1);
        mbar_init_fence();
    }
    __syncwarp();
    uint32_t boxes_done = 0;  // boxes consumed by this warp so far: fixes ring slot and mbarrier phase

    int strip, chunk, env;
    long long unit_id;
    while (next_unit(p, par, lane, strip, chunk, env, unit_id)) {
        if (!p.meta[(long long)par * p.meta_stride + env].running) continue;
        uint32_t emitted = 0;
        const int x0 = strip * WR;
        const int y_begin = chunk * p.rows_per_chunk;
        const int n_rows = min(y_begin + p.rows_per_chunk, p.H) - y_begin;  // rows this unit owns
        const int n_box = (n_rows + 2 + B - 1) / B;      // local row j <-> grid row y_begin - 1 + j
        const int c0 = (x0 - CPL) * (int)sizeof(CellT) / 4;
        const uint32_t kb = boxes_done;                  // global index of this unit's box 0

        auto issue_box = [&](int k) {  // lane 0 only
            const uint32_t st = (kb + k) % TMA_STAGES;
            mbar_expect_tx(bar0 + 8 * st, TMA_BOX_BYTES);
            tma_load_3d(ring_u32 + st * TMA_BOX_BYTES, &tmap, c0, y_begin - 1 + k * B, env, bar0 + 8 * st);
        };
        __syncwarp();  // every lane is done with the previous unit's boxes
        if (lane == 0)
            for (int k = 0; k < min(n_box, TMA_STAGES); ++k) issue_box(k);

        // bit i of `nz`: local row (k*B - 2 + i) holds something to look at; two rows carried over
        uint32_t carry = 0;
        for (int k = 0; k < n_box; ++k) {
            const uint32_t K = kb + k;
            mbar_wait(bar0 + 8 * (K % TMA_STAGES), (K / TMA_STAGES) & 1);
            const CellT* box = ring + (K % TMA_STAGES) * (B * RS);
            uint32_t nz = carry;
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const uint4 v = *reinterpret_cast<const uint4*>(box + i * RS + CPL + lane * CPL);
                const bool a = (((v.x | v.y) | (v.z | v.w)) & look_mask) != 0;
                if (__any_sync(0xffffffffu, a)) nz |= 4u << i;
            }
            if (p.strips > 1) {  // cells just outside the strip: lanes 0..B-1 the left pad of row `lane`, B..2B-1 the right pad
                uint32_t hw = 0;
                if (lane < 2 * B) {
                    const unsigned char* rowb = reinterpret_cast<const unsigned char*>(box + (lane & (B - 1)) * RS);
                    hw = *reinterpret_cast<const uint32_t*>(rowb + (lane < B ? 12 : TMA_ROW_BYTES - 16));
                    // only the adjacent cell counts: the last cell of the left pad, the first of the right pad
                    hw = lane < B ? (hw >> (32 - 8 * (int)sizeof(CellT))) : (hw & CELL_ALL);
                }
                const uint32_t hb = __ballot_sync(0xffffffffu, (hw & look_mask & CELL_ALL) != 0);
                nz |= ((hb | (hb >> B)) & ((1u << B) - 1u)) << 2;
            }
            carry = nz >> B;
            // the box is consumed (only `nz` survives): refill its slot right away
            __syncwarp();
            if (lane == 0 && k + TMA_STAGES < n_box) issue_box(k + TMA_STAGES);
            // local row j = k*B - 1 + i is decided once box k is here (its row j+1 = k*B + i), i = 0..B-1
            const uint32_t rows = (nz | (nz >> 1) | (nz >> 2)) & ((1u << B) - 1u);
            if (rows) {  // warp-uniform
                const int j = k * B - 1 + lane;
                emitted |= tasks.push(lane < B && ((rows >> lane) & 1u) && j >= 1 && j <= n_rows, make_row_task(env, y_begin - 1 + j, strip));
            }
        }
        boxes_done += (uint32_t)n_box;
        retire_unit(p, lane, unit_id, emitted);
    }
    tasks.flush();
}

// ---- front end 2: 128-bit loads ------------------------------------------------------------
template <typename CellT>
__global__ void __launch_bounds__(SWEEP_WARPS * 32, SFB_LDG_MIN_BLOCKS) k_sweep_ldg(const DevParams p, const int par) {
    using C = Cell<CellT>;
    constexpr int CPL = C::CPL, WR = 32 * CPL;
    constexpr uint32_t CELL_ALL = sizeof(CellT) == 1 ? 0xFFu : 0xFFFFu;
    __shared__ unsigned long long tl_all[SWEEP_WARPS][WQ_CAP];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = p.H, pitch = p.pitch;
    const CellT* const filler = reinterpret_cast<const CellT*>(p.filler) + CPL;  // valid for [-CPL, pitch + CPL)
    RowTaskList tasks(p, par, lane, tl_all[warp]);
    const uint32_t look_mask = p.attenuate ? (C::CODE_MASK | C::LINE_MASK) : C::CODE_MASK;

    int strip, chunk, env;
    long long unit_id;
    while (next_unit(p, par, lane, strip, chunk, env, unit_id)) {
        if (!p.meta[(long long)par * p.meta_stride + env].running) continue;
        uint32_t emitted = 0;
        const CellT* const envbase = reinterpret_cast<const CellT*>(p.state) + (long long)env * p.plane;
        const int x0 = strip * WR;
        const int xl = x0 + lane * CPL;
        const int y_begin = chunk * p.rows_per_chunk;
        const int y_end = min(y_begin + p.rows_per_chunk, H);
        const bool in_x = xl < pitch;
        // lanes 0 / 31 fetch the cell just outside the strip (if there is one)
        const int hoff = lane == 0 ? x0 - 1 : x0 + WR;
        const bool hpred = (lane == 0 && strip > 0) || (lane == 31 && x0 + WR < pitch);

        // rows outside [0, H): the neighbour slab's edge row in slab mode, BURNED filler otherwise
        auto edge_row = [&](int y) -> const CellT* {
            if (y == -1 && p.halo_top) return reinterpret_cast<const CellT*>(p.halo_top) + (long long)env * p.halo_top_plane;
            if (y == H && p.halo_bottom) return reinterpret_cast<const CellT*>(p.halo_bottom) + (long long)env * p.halo_bottom_plane;
            return filler;
        };
        struct Row { uint4 v; uint32_t h; };
        auto issue = [&](const CellT* rowp, Row& r) {
            r.v = make_uint4(C::FILL, C::FILL, C::FILL, C::FILL);
            r.h = ST_BURNED;
            if (in_x) r.v = *reinterpret_cast<const uint4*>(rowp + xl);
            if (hpred) r.h = rowp[hoff];
        };
        auto busy = [&](const Row& r) -> bool {  // warp-uniform: does the row hold anything to look at?
            const uint32_t a = (((r.v.x | r.v.y) | (r.v.z | r.v.w)) & look_mask) | (r.h & look_mask & CELL_ALL);
            return __any_sync(0xffffffffu, a != 0);
        };

        // rolling window of "busy" flags; the four loads of a batch are issued before any is used
        Row r2, r3, r4, r5;
        const CellT* rowp = envbase + (long long)y_begin * pitch;  // row y of the loop below
        issue(y_begin > 0 ? rowp - pitch : edge_row(-1), r2);
        issue(rowp, r3);
        bool b0 = busy(r2), b1 = busy(r3);
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
            const bool b2 = busy(r2), b3 = busy(r3), b4 = busy(r4), b5 = busy(r5);
            // rows y .. y+3: lane i decides row y + i
            const uint32_t rows = ((b0 | b1 | b2) ? 1u : 0u) | ((b1 | b2 | b3) ? 2u : 0u) | ((b2 | b3 | b4) ? 4u : 0u) |
                                  ((b3 | b4 | b5) ? 8u : 0u);
            if (rows) emitted |= tasks.push(lane < 4 && ((rows >> lane) & 1u) && y + lane < y_end, make_row_task(env, y + lane, strip));
            b0 = b4;
            b1 = b5;
            rowp = q4;
        }
        retire_unit(p, lane, unit_id, emitted);
    }
    tasks.flush();
}

// ---------------------------------------------------------------------------------------
// k_rows: one warp per row task.
// ---------------------------------------------------------------------------------------
#ifndef SFB_ROWS_CHUNK
#define SFB_ROWS_CHUNK 4
#endif
constexpr int ROWS_WARPS = 4;
constexpr int ROWS_CHUNK = SFB_ROWS_CHUNK;

template <typename CellT>
__global__ void __launch_bounds__(ROWS_WARPS * 32) k_rows(const DevParams p, const int par) {
    using RW = RowWorker<CellT>;
    using C = Cell<CellT>;
    constexpr int CPL = RW::CPL, WR = RW::WR, RS = RW::RS;
    __shared__ __align__(16) CellT sm_all[ROWS_WARPS][3][RS];
    __shared__ unsigned long long wq_all[ROWS_WARPS][WQ_CAP];  // per-warp staging of work items
    __shared__ uint32_t lut_all[ROWS_WARPS][64];                // seg_lut | key_lut

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned long long n = min(p.rows_count[par], (unsigned long long)p.rows_cap);
    const unsigned long long n_warps = (unsigned long long)gridDim.x * ROWS_WARPS;
    const unsigned long long w = (unsigned long long)blockIdx.x * ROWS_WARPS + warp;
    // chunk of consecutive tasks per deal: small lists are spread task by task
    const unsigned long long chunk = n >= n_warps * ROWS_CHUNK ? ROWS_CHUNK : (n >= n_warps * 2 ? 2 : 1);
    if (w * chunk >= n) return;

    RW::build_seg_lut(lut_all[warp], lane);
    RW rw(p, par, lane, wq_all[warp], lut_all[warp]);
#ifdef SFB_ROWS_REDUX
    rw.my_groups = lut_all[warp][lane];
#endif
    CellT(*sm)[RS] = sm_all[warp];
    const int H = p.H, pitch = p.pitch;
    const long long plane = p.plane;
    const CellT* const state = reinterpret_cast<const CellT*>(p.state);
    const CellT* const filler = reinterpret_cast<const CellT*>(p.filler) + CPL;
    // rows outside [0, H): the neighbour slab's edge row in slab mode, BURNED filler otherwise
    auto edge_row = [&](int yy, int env) -> const CellT* {
        if (yy == -1 && p.halo_top) return reinterpret_cast<const CellT*>(p.halo_top) + (long long)env * p.halo_top_plane;
        if (yy == H && p.halo_bottom) return reinterpret_cast<const CellT*>(p.halo_bottom) + (long long)env * p.halo_bottom_plane;
        return filler;
    };

    // chunks of ROWS_CHUNK consecutive tasks, dealt round-robin to the warps: consecutive tasks are
    // mostly consecutive rows of one env (two of the three rows of a task then come from L1),
    // and dealing small chunks spreads a long fire front over many warps
    for (unsigned long long c = w; c * chunk < n; c += n_warps) {
        const unsigned long long t_end = min((c + 1) * chunk, n);
        for (unsigned long long t = c * chunk; t < t_end; ++t) {
            const unsigned long long task = p.rows[t];
            const int y = (int)(task & 0xFFFFFu), strip = (int)((task >> 20) & 0xFFu), env = (int)(task >> 28);
            if (env != rw.env) rw.set_env(env, p.meta[(long long)par * p.meta_stride + env]);
#ifdef SFB_ROWS_PADVEC
            // variant: the env's base and edge rows come from set_env; lanes 0 / 31 fetch the 16 bytes left /
            // right of the strip as one vector each and store it into the staged row's pad
            const int x0 = strip * WR;
            rw.x0 = x0;
            const int xl = x0 + lane * CPL;
            const bool in_x = xl < pitch;
            const bool edge_lane = lane == 0 || lane == 31;
            const bool hpred = (lane == 0 && strip > 0) || (lane == 31 && x0 + WR < pitch);
            const int poff = lane == 0 ? x0 - CPL : x0 + WR;  // first cell of the pad vector in the row
            const int pslot = lane == 0 ? 0 : CPL + WR;       // ... and in the staged row
            const CellT* rowp[3];
            rowp[1] = rw.envbase + (long long)y * pitch;
            rowp[0] = y > 0 ? rowp[1] - pitch : rw.row_above;
            rowp[2] = y + 1 < H ? rowp[1] + pitch : rw.row_below;
            uint4 v[3], vp[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                v[r] = make_uint4(C::FILL, C::FILL, C::FILL, C::FILL);
                vp[r] = v[r];
                if (in_x) v[r] = *reinterpret_cast<const uint4*>(rowp[r] + xl);
                if (hpred) vp[r] = *reinterpret_cast<const uint4*>(rowp[r] + poff);
            }
            __syncwarp();  // the previous task's readers are done with the staging rows
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                *reinterpret_cast<uint4*>(&sm[r][CPL + lane * CPL]) = v[r];
                if (edge_lane) *reinterpret_cast<uint4*>(&sm[r][pslot]) = vp[r];
            }
            __syncwarp();
            const uint4 vo = make_uint4(v[0].x | v[1].x | v[2].x, v[0].y | v[1].y | v[2].y, v[0].z | v[1].z | v[2].z,
                                        v[0].w | v[1].w | v[2].w);
            // the one cell that touches the strip: the last of the left pad, the first of the right pad
            const uint32_t hv = lane == 0 ? ((vp[0].w | vp[1].w | vp[2].w) >> (32 - 8 * (int)sizeof(CellT)))
                                          : ((vp[0].x | vp[1].x | vp[2].x) & RW::CELL_ALL);
            const uint32_t act = __ballot_sync(0xffffffffu, rw.seg_needs_look(vo, edge_lane ? hv : 0u));
#else
            const int x0 = strip * WR;
            rw.x0 = x0;
            const int xl = x0 + lane * CPL;
            const bool in_x = xl < pitch;
            const int hoff = lane == 0 ? x0 - 1 : x0 + WR;
            const bool hpred = (lane == 0 && strip > 0) || (lane == 31 && x0 + WR < pitch);
            const CellT* const envbase = state + (long long)env * plane;
            // rows y-1, y, y+1: one 64-bit multiply; only the grid's first / last row takes the slow path
            const CellT* rowp[3];
            rowp[1] = envbase + (long long)y * pitch;
            rowp[0] = y > 0 ? rowp[1] - pitch : edge_row(-1, env);
            rowp[2] = y + 1 < H ? rowp[1] + pitch : edge_row(H, env);
            uint4 v[3];
            uint32_t h[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                v[r] = make_uint4(C::FILL, C::FILL, C::FILL, C::FILL);
                h[r] = ST_BURNED;
                if (in_x) v[r] = *reinterpret_cast<const uint4*>(rowp[r] + xl);
                if (hpred) h[r] = rowp[r][hoff];
            }
            __syncwarp();  // the previous task's readers are done with the staging rows
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                *reinterpret_cast<uint4*>(&sm[r][CPL + lane * CPL]) = v[r];
                if (lane == 0) sm[r][CPL - 1] = (CellT)h[r];
                if (lane == 31) sm[r][CPL + WR] = (CellT)h[r];
            }
            __syncwarp();
            const uint4 vo = make_uint4(v[0].x | v[1].x | v[2].x, v[0].y | v[1].y | v[2].y, v[0].z | v[1].z | v[2].z,
                                        v[0].w | v[1].w | v[2].w);
            const uint32_t act = __ballot_sync(0xffffffffu, rw.seg_needs_look(vo, h[0] | h[1] | h[2]));
#endif
            // row units: nothing to look at in the 3-row window -> the row leaves the list until an
            // ignition, a control line or a map upload next to it flags it again
            if (p.unit_rows && act == 0 && lane == 0) p.unit_act[(long long)env * p.unit_stride + (long long)y * p.strips + strip] = 0;
            rw.detail_row(y, sm[0], sm[1], sm[2], act);
        }
    }
    rw.finish();
}

// ---------------------------------------------------------------------------------------
// k_eval: persistent grid-stride over the work queue (or, if the queue overflowed, over
// every cell).  Threads 0..E-1 also write the next step's EnvMeta.
// ---------------------------------------------------------------------------------------
// the per-env clock of the next step from this step's flags (fire.py:633-652, :717).  The flags were raised by
// other blocks (of an earlier kernel, or -- bitboard handles without a k_eval -- of this one): volatile reads
__device__ __forceinline__ void advance_env(const DevParams& p, const int par, const long long env) {
    // two 16-byte loads that bypass L1 (this SM may hold the line from before the flags were raised) and, unlike
    // volatile loads, may be in flight together with those of the caller's other envs
    const int4* c = reinterpret_cast<const int4*>(p.meta + (long long)par * p.meta_stride + env);
#ifdef SFB_EMU
    const int4 lo = c[0], hi = c[1];
#else
    const int4 lo = __ldcg(c), hi = __ldcg(c + 1);
#endif
    EnvMeta nxt;
    nxt.t = lo.x;
    nxt.running = lo.y;
    const int time_quit = lo.z, any_live = lo.w, any_cand = hi.x;
    nxt.pad = 0;
    {  // bytes 24..31 of the record
        const unsigned long long bits = (unsigned long long)(uint32_t)hi.z | ((unsigned long long)(uint32_t)hi.w << 32);
        memcpy(&nxt.elapsed, &bits, sizeof(double));
    }
    if (nxt.running) {
        if (!any_live) nxt.running = 0;            // fire.py:637
        else if (time_quit) nxt.running = 0;       // fire.py:641-643
        else if (any_cand) nxt.elapsed = nxt.elapsed + p.dt;  // fire.py:717 (skipped by :651)
        nxt.t = nxt.t + 1;
    }
    nxt.any_live = 0;
    nxt.any_cand = 0;
    nxt.time_quit = p.has_max_time && (p.dt > p.max_time || nxt.elapsed > p.max_time);
    p.meta[(long long)(par ^ 1) * p.meta_stride + env] = nxt;
}
// one thread, after the step's last work: counters of the next step
__device__ __forceinline__ void close_step(const DevParams& p, const int par) {
    p.qcount[par ^ 1] = 0;
    p.overflow[par ^ 1] = 0;
    p.unit_next[par ^ 1] = 0;
    p.rows_next[par ^ 1] = 0;
    if (p.bits) {  // the tile list of the next step is being written: the one just consumed is emptied
        p.units_count[par] = p.rows_count[par];  // (kept for sfb_get_unit_stats)
        p.rows_count[par] = 0;
    } else {
        p.rows_count[par ^ 1] = 0;
        if (p.units_count) p.units_count[par ^ 1] = 0;
    }
}

// bitboard handles (sfb_bits.cuh)
template <typename CellT>
__device__ __forceinline__ void bits_unring(const DevParams& p, const EnvMeta& m, int env, long long idx, int slot);
template <typename CellT>
__device__ __forceinline__ void bits_unring_all(const DevParams& p, const EnvMeta& m, int env, int y, int x);
constexpr int DIR_UNRING = 10;  // work item of a bitboard handle: the ring bit of a sprite replaced by a re-ignition

template <typename CellT>
__device__ void dense_cell(const DevParams& p, const int par, long long idx) {
    using C = Cell<CellT>;
    const int env = (int)(idx / p.plane);
    const long long cell = idx - (long long)env * p.plane;
    const int y = (int)(cell / p.pitch), x = (int)(cell - (long long)y * p.pitch);
    if (x >= p.W) return;
    const EnvMeta m = p.meta[(long long)par * p.meta_stride + env];
    if (!m.running || m.time_quit) return;
    const CellT* st = reinterpret_cast<const CellT*>(p.state);
    const int s = st[idx] & 7;
    // bitboard handles: k_tiles has evaluated the candidates; the items it left for k_eval were lost with the queue
    if (p.bits && ((int)st[idx] >> 3) == 1 + (m.t % C::M)) bits_unring_all<CellT>(p, m, env, y, x);
    if (!ignitable(s)) return;
    const int tm1 = (m.t - 1) % C::M;
    int best = p.max_dur, dir = DIR_NONE;
    auto look = [&](int dy, int dx, int k) {
        const int yy = y + dy, xx = x + dx;
        if (xx < 0 || xx >= p.W) return;
        const CellT* rowp;
        if (yy >= 0 && yy < p.H) rowp = st + (long long)env * p.plane + (long long)yy * p.pitch;
        else if (yy == -1 && p.halo_top) rowp = reinterpret_cast<const CellT*>(p.halo_top) + (long long)env * p.halo_top_plane;
        else if (yy == p.H && p.halo_bottom) rowp = reinterpret_cast<const CellT*>(p.halo_bottom) + (long long)env * p.halo_bottom_plane;
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
    if (p.bits && dir != DIR_NONE) return;  // a candidate: done by k_tiles
    process_item<CellT>(p, m, env, idx, dir, s);
}

template <typename CellT>
__global__ void __launch_bounds__(256) k_eval(const DevParams p, const int par) {
    grid_dep_wait();
    grid_dep_launch();
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gstride = (long long)gridDim.x * blockDim.x;

    if (p.overflow[par]) {
        const long long total = (long long)p.E * p.plane;
        for (long long i = gid; i < total; i += gstride) dense_cell<CellT>(p, par, i);
        if (p.track && gid == 0) *p.chg_overflow = 1;  // the dense pass does not log: force a full resync
    } else {
        const long long n = (long long)p.qcount[par];
        const int lane = threadIdx.x & 31;
        for (long long base = gid - lane; base < n; base += gstride) {  // warp-uniform trip count
            const long long i = base + lane;
            long long idx = 0;
            int logged = -1;
            if (i < n) {
                const unsigned long long it = p.queue[i];
                idx = (long long)(it & 0xFFFFFFFFFFFFull);
                const int dir = (int)((it >> 48) & 0xF), s = (int)((it >> 52) & 7);
                if (dir == DIR_PRUNED) {
                    logged = 2;  // BurnStatus.BURNED
                } else if (dir == DIR_UNRING) {
                    const int env = (int)(idx / p.plane);
                    bits_unring<CellT>(p, p.meta[(long long)par * p.meta_stride + env], env, idx, s);
                } else {
                    const int env = (int)(idx / p.plane);
                    const EnvMeta m = p.meta[(long long)par * p.meta_stride + env];
                    if (process_item<CellT>(p, m, env, idx, dir, s)) logged = 1;  // BurnStatus.BURNING
                }
            }
            if (p.track) log_append(p, logged >= 0, idx + p.idx_base, logged);
        }
    }

    // per-env clock for the next step (reads only what k_sweep finalised)
    for (long long env = gid; env < p.E; env += gstride) advance_env(p, par, env);
    if (gid == 0) close_step(p, par);
}

// ---------------------------------------------------------------------------------------
// slab mode synchronisation kernels (one warp each)
// ---------------------------------------------------------------------------------------
constexpr long long SLAB_SPIN_LIMIT = 1ll << 27;  // ~ a second of polling, then give up instead of hanging

// after k_sweep of global step `gstep`: publish this slab's any_live / any_cand to every slab,
// wait for everybody's, OR them into this slab's EnvMeta
__global__ void k_slab_exchange_flags(const DevParams p, const int par, const uint32_t gstep) {
    const int lane = threadIdx.x;
    const int half = gstep & 1;
    EnvMeta* meta = p.meta + (long long)par * p.meta_stride;
    for (int q = 0; q < p.slab_world; ++q) {
        SlabMailbox* box = p.peer_box[q];
        for (int i = lane; i < p.E * 2; i += 32) {
            const EnvMeta& m = meta[i >> 1];
            box->flags[half][p.slab_rank][i >> 1][i & 1] = (i & 1) ? m.any_cand : m.any_live;
        }
    }
    __threadfence_system();
    __syncwarp();
    if (lane < p.slab_world) p.peer_box[lane]->flag_step[half][p.slab_rank] = gstep;
    __threadfence_system();
    bool ok = true;
    if (lane < p.slab_world) {
        long long spins = 0;
        while (p.mailbox->flag_step[half][lane] != gstep) {
            if (++spins > SLAB_SPIN_LIMIT) { ok = false; break; }
        }
    }
    __threadfence_system();
    if (!__all_sync(0xffffffffu, ok)) {
        if (lane == 0) p.mailbox->error = 1;
        return;
    }
    for (int i = lane; i < p.E * 2; i += 32) {
        int v = 0;
        for (int q = 0; q < p.slab_world; ++q) v |= p.mailbox->flags[half][q][i >> 1][i & 1];
        if (v) {
            if (i & 1) meta[i >> 1].any_cand = 1;
            else meta[i >> 1].any_live = 1;
        }
    }
}

// after k_eval of global step `gstep`: tell the neighbours, and wait until they have finished it too
// (their ignitions of this step must be visible before the next sweep reads their edge rows)
__global__ void k_slab_step_done(const DevParams p, const uint32_t gstep) {
    const int lane = threadIdx.x;
    const int up = p.slab_rank - 1, down = p.slab_rank + 1;
    __threadfence_system();
    if (lane == 0 && up >= 0) p.peer_box[up]->done_step[p.slab_rank] = gstep;
    if (lane == 1 && down < p.slab_world) p.peer_box[down]->done_step[p.slab_rank] = gstep;
    __threadfence_system();
    bool ok = true;
    const int from = lane == 0 ? up : (lane == 1 ? down : -1);
    if (from >= 0 && from < p.slab_world) {
        long long spins = 0;
        while ((int32_t)(p.mailbox->done_step[from] - gstep) < 0) {
            if (++spins > SLAB_SPIN_LIMIT) { ok = false; break; }
        }
    }
    __threadfence_system();
    if (!__all_sync(0xffffffffu, ok) && lane == 0) p.mailbox->error = 1;
}

}  // namespace sfb
