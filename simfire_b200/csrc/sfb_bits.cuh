// Bitboard front end (SFB_FRONT_BITS): one update() call as word-wide bit operations on tiles of
// 32 rows x 30 columns, with the rate look-up, the burn accumulation and the ignition test of the
// tile's candidates done by the same warp -- instead of a cell-by-cell look at 512-cell rows (k_rows)
// followed by a pass over a global work queue (k_eval).
//
// Next to the canonical per-cell state bytes the handle keeps bit planes, one 32-bit word per
// (env, tile column, row).  A word owns 30 cells -- bit 1 + (x mod 30) of word x / 30 -- and, in the
// sprite planes, carries a copy of the cell left of them in bit 0 and of the cell right of them in bit 31, so
// that the three rows y-1, y, y+1 of ONE word column hold the whole 3 x 3 neighbourhood of its 30 cells:
//
//   IGN    the cell is ignitable (UNBURNED or a control line, fire.py:192-205); owned bits only
//   LINE   the cell is a control line; owned bits only
//   RING   R = max_fire_duration + 1 planes: plane (s mod R) holds the Fire sprites created by update()
//          call s.  In call t the sprites of duration a (fire.py:633) are plane (t - 1 - a) mod R: the
//          sources of the step are R - 1 of the planes as they stand; the remaining plane (t mod R)
//          holds the sprites that have reached max_fire_duration -- they are pruned (fire.py:116-161)
//          and the cells ignited by this call are written into it.  No plane that is read as a source
//          is written during the step, every bit of a word has exactly one writer (the owner of the cell;
//          all updates are atomic bit operations), and nothing but the owner of a cell touches its state
//          byte or burn value: tiles never race, whatever the order the warps take them in.
//
// A step is k_tiles + k_eval: the list of tiles a step has to look at is written by the step before it
// (and by whatever touches the maps between steps), so nothing is scanned to find them.
//
// k_tiles, one warp per listed tile, lane = row:
//   * prune: the owned bits of plane t mod R -> BURNED state bytes, bits cleared (also the copies);
//   * candidate search: for every duration (youngest first) and every neighbour in the reference's
//     last-write-wins order (S-E, S, S-W, E, W, N-E, N, N-W: fire.py:704-705 + sprite-list order) the shifted
//     source word is ANDed with the still undecided ignitable cells; three more words collect the winning
//     direction.  ~25 warp-instructions per duration decide the 960 cells of a tile;
//   * the candidates are compacted into shared memory and evaluated one per lane (process_item: the
//     tabulated rate of the (cell, direction) pair, attenuation, float64 burn +=, strict > pixel_scale);
//     the cells that ignite are collected per row and written to plane t mod R with one atomic per word;
//   * control-line cells no fire touches are attenuated only if the env has any candidate in this call
//     (fire.py:271-278, :651-652): a whole-env fact, so they go to the work queue and k_eval, which runs
//     after every tile, applies them; k_eval also advances the per-env clocks, as in every front end.
//
// A tile is listed while the 34 x 32 window of its words holds a sprite bit or (with attenuation) the
// tile holds a control line.  Lists and their flags (tile_act: "this tile is on the list", so that nobody
// lists it twice) are double-buffered by step parity: a step takes the tiles of its parity off their list
// and puts on the next step's list the tile itself if it still has something to look at and the neighbours of
// a cell that ignites on the tile's border, so taking off and putting on never meet on one byte.  Resets,
// mitigation and map uploads list tiles for the step that comes next.  Bitboard handles step all their envs
// as one group (a step is two short kernels; there is nothing to overlap).
#pragma once
#include "sfb_kernels.cuh"

namespace sfb {

constexpr int BP_IGN = 0, BP_LINE = 1, BP_RING = 2;
constexpr int BITS_MAX_DUR = 7;  // ring of at most 8 planes; longer-lived sprites use the byte front ends
constexpr int TW = 30;           // cells a word owns
constexpr uint32_t OWN = 0x7FFFFFFEu;

__device__ __forceinline__ uint32_t* bits_word(const DevParams& p, int env, int plane, int tx, int y) {
    return p.bits + (long long)env * p.bits_env + (long long)plane * p.bits_plane + (long long)tx * p.H + y;
}
// plane of the sprites that have duration `a` in update() call t (1-based)
__device__ __forceinline__ int ring_slot(int t, int a, int R) {
    int s = (t - 1 - a) % R;
    return s < 0 ? s + R : s;
}
// "tile is on list `buf`": one byte per tile, set with an atomic OR on the word that holds it; true if this
// call set it (the caller then appends the tile)
__device__ __forceinline__ bool tile_set_atomic(const DevParams& p, int buf, long long tile) {
    uint8_t* const f = p.tile_act + (long long)buf * p.tile_buf;
    const uint32_t bit = 1u << (8 * (int)(tile & 3));
    return (atomicOr(reinterpret_cast<uint32_t*>(f) + (tile >> 2), bit) & bit) == 0;
}
__device__ __forceinline__ bool tile_test_and_set(const DevParams& p, int buf, long long tile) {
    if (*reinterpret_cast<volatile uint8_t*>(p.tile_act + (long long)buf * p.tile_buf + tile)) return false;  // listed already
    return tile_set_atomic(p, buf, tile);
}
// tile task: ty | tx << 16 | env << 32
__device__ __forceinline__ unsigned long long make_tile_task(int env, int ty, int tx) {
    return (unsigned long long)(unsigned)ty | ((unsigned long long)(unsigned)tx << 16) | ((unsigned long long)(unsigned)env << 32);
}
// put tile (ty, tx) of `env` on list `buf` (setup kernels; k_tiles stages its appends per block)
__device__ __forceinline__ void tile_list_add(const DevParams& p, int buf, int env, int ty, int tx) {
    if (!tile_test_and_set(p, buf, (long long)env * p.tile_stride + (long long)ty * p.tiles_x + tx)) return;
    const unsigned long long slot = atomicAdd(p.rows_count + buf, 1ULL);
    if (slot < (unsigned long long)p.rows_cap) p.rows[(long long)buf * p.rows_cap + slot] = make_tile_task(env, ty, tx);
}
// set / clear the bit of cell (y, x) in a sprite plane: the owned bit and its copies in the neighbour words
__device__ __forceinline__ void ring_set_cell(const DevParams& p, int env, int slot, int y, int x, bool on) {
    const int tx = x / TW, b = x - tx * TW;
    uint32_t* w = bits_word(p, env, BP_RING + slot, tx, y);
    if (on) {
        atomicOr(w, 2u << b);
        if (b == 0 && tx > 0) atomicOr(w - p.H, 1u << 31);
        if (b == TW - 1 && tx + 1 < p.tiles_x) atomicOr(w + p.H, 1u);
    } else {
        atomicAnd(w, ~(2u << b));
        if (b == 0 && tx > 0) atomicAnd(w - p.H, ~(1u << 31));
        if (b == TW - 1 && tx + 1 < p.tiles_x) atomicAnd(w + p.H, ~1u);
    }
}
// fire_map[y, x] was set to internal status `s` from outside (mitigation.py:77): sprite planes untouched
__device__ __forceinline__ void bits_on_status(const DevParams& p, int env, int y, int x, int s) {
    const int tx = x / TW;
    const uint32_t bit = 2u << (x - tx * TW);
    uint32_t* ign = bits_word(p, env, BP_IGN, tx, y);
    uint32_t* line = bits_word(p, env, BP_LINE, tx, y);
    if (ignitable(s)) atomicOr(ign, bit);
    else atomicAnd(ign, ~bit);
    if (s & ST_LINE_BIT) {
        atomicOr(line, bit);
        if (p.attenuate) tile_list_add(p, p.bits_par, env, y >> 5, tx);
    } else {
        atomicAnd(line, ~bit);
    }
}
// k_eval, DIR_UNRING item: the cell carried a live sprite of plane `slot` when it was a candidate; if
// it re-ignited in this call (its byte carries this call's code) the newer sprite replaces the older one
template <typename CellT>
__device__ __forceinline__ void bits_unring(const DevParams& p, const EnvMeta& m, int env, long long idx, int slot) {
    const int c = (int)reinterpret_cast<const CellT*>(p.state)[idx] >> 3;
    if (c != 1 + (m.t % Cell<CellT>::M)) return;
    const long long cell = idx - (long long)env * p.plane;
    const int y = (int)(cell / p.pitch), x = (int)(cell - (long long)y * p.pitch);
    ring_set_cell(p, env, slot, y, x, false);
}

// the dense form of the same (k_eval after a queue overflow): a cell that ignited in this call keeps no bit in
// any source plane
template <typename CellT>
__device__ __forceinline__ void bits_unring_all(const DevParams& p, const EnvMeta& m, int env, int y, int x) {
    const int e = m.t % p.ring;
    for (int slot = 0; slot < p.ring; ++slot)
        if (slot != e) ring_set_cell(p, env, slot, y, x, false);
}

// (re)derive every plane of n envs (a device list, or [env0, env0 + n)) from their state bytes; one
// thread per word.  Durations are those the NEXT update() call will see; flags are raised for that call.
template <typename CellT>
__global__ void k_bits_rebuild(const DevParams p, const int par, const int32_t* envs, const int env0, const int n) {
    const long long per_env = (long long)p.tiles_x * p.H, total = (long long)n * per_env;
    const CellT* state = reinterpret_cast<const CellT*>(p.state);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / per_env);
        const int env = envs ? envs[k] : env0 + k;
        const long long r = i - (long long)k * per_env;
        const int y = (int)(r % p.H), tx = (int)(r / p.H);  // consecutive threads: consecutive rows of one word column
        const int t = p.meta[(long long)par * p.meta_stride + env].t;
        const int tm1 = (t - 1) % Cell<CellT>::M;
        const CellT* row = state + (long long)env * p.plane + (long long)y * p.pitch;
        uint32_t ign = 0, line = 0, ring[BITS_MAX_DUR + 1];
        for (int s = 0; s <= BITS_MAX_DUR; ++s) ring[s] = 0;
        for (int b = 0; b < 32; ++b) {
            const int x = tx * TW - 1 + b;
            if (x < 0 || x >= p.W) continue;
            const int c = row[x];
            const int st = c & 7;
            const bool own = b >= 1 && b <= TW;
            if (own && ignitable(st)) ign |= 1u << b;
            if (own && (st & ST_LINE_BIT)) line |= 1u << b;
            if ((c >> 3) != 0) {
                const int a = min(sprite_age<CellT>(c >> 3, tm1), p.max_dur);  // >= max_dur: pruned by the next call
                ring[ring_slot(t, a, p.ring)] |= 1u << b;
            }
        }
        *bits_word(p, env, BP_IGN, tx, y) = ign;
        *bits_word(p, env, BP_LINE, tx, y) = line;
        uint32_t any = (p.attenuate ? line : 0u);
        for (int s = 0; s < p.ring; ++s) {
            *bits_word(p, env, BP_RING + s, tx, y) = ring[s];
            any |= ring[s];
        }
        if (any) {  // every tile whose window can see a bit of this word (generously)
            const int ty0 = (y > 0 ? y - 1 : 0) >> 5, ty1 = (y + 1 < p.H ? y + 1 : p.H - 1) >> 5;
            for (int ty = ty0; ty <= ty1; ++ty)
                for (int txx = max(tx - 1, 0); txx <= min(tx + 1, p.tiles_x - 1); ++txx) tile_list_add(p, par, env, ty, txx);
        }
    }
}

#ifndef SFB_TILES_WARPS
#define SFB_TILES_WARPS 4
#endif
#ifndef SFB_TILES_MIN_BLOCKS
#define SFB_TILES_MIN_BLOCKS 8
#endif
constexpr int TILES_WARPS = SFB_TILES_WARPS;
constexpr int AQ_CAP = 32 * TILES_WARPS;

// NSRC = max_fire_duration = the number of sprite planes a step reads as sources (the ring has NSRC + 1)
// fuse != 0 (SFB_FUSE_EVAL=1, off by default: measured slower than the second launch it saves): there is no
// k_eval behind this launch -- the block that finishes last (atomic ticket) closes the step: whatever is in the
// work queue, the per-env clocks, the counters.  The host only asks for it while nothing has written a status from
// outside since the envs were reset (no control lines, so the queue stays empty and closing the step is E clock
// updates); with control lines the queue can be long and k_eval's whole grid takes it.
template <typename CellT, int NSRC, bool STATS>
__global__ void __launch_bounds__(TILES_WARPS * 32, SFB_TILES_MIN_BLOCKS) k_tiles(const DevParams p, const int par, const int fuse) {
    using C = Cell<CellT>;
    grid_dep_wait();
    grid_dep_launch();
    __shared__ uint16_t cq_all[TILES_WARPS][WQ_CAP];            // candidates of the tile in hand: row | bit << 5
    __shared__ unsigned long long dq_all[TILES_WARPS][WQ_CAP];  // staging of items for k_eval (deferred line cells, replaced sprites)
    __shared__ uint32_t nb_all[TILES_WARPS][32];                // cells of the tile that ignited, per row
    __shared__ unsigned long long lq_all[TILES_WARPS][WQ_CAP];  // staging of change-log entries (SFB_TRACK_CHANGES)
    __shared__ unsigned long long pq_all[TILES_WARPS][32];      // per lane: the tile it may have listed (resolve)
    __shared__ unsigned long long aq[AQ_CAP];                   // the block's appends to the next step's tile list
    __shared__ unsigned int aq_n, aq_base;
    if (threadIdx.x == 0) aq_n = 0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    uint16_t* const cq = cq_all[warp];
    unsigned long long* const dq = dq_all[warp];
    uint32_t* const nb = nb_all[warp];
    unsigned long long* const lq = lq_all[warp];
    int dcount = 0, lcount = 0;  // warp-uniform
    unsigned int st_tiles = 0, st_cand = 0, st_ign = 0, st_pruned = 0, st_def = 0;  // statistics (kernel-timing passes only)
    const unsigned long long* const tasks = p.rows + (long long)par * p.rows_cap;
    // the warp's first task is fetched together with the length of the list, not after it (stale or not, the
    // slot exists; it is only used if the list is that long)
    const unsigned long long ti0 = (unsigned long long)blockIdx.x * TILES_WARPS + (threadIdx.x >> 5);
    unsigned long long task = tasks[min(ti0, (unsigned long long)p.rows_cap - 1)];
    const unsigned long long n = min(p.rows_count[par], (unsigned long long)p.rows_cap);
    unsigned long long* const tasks_nxt = p.rows + (long long)(par ^ 1) * p.rows_cap;
    uint8_t* const flags_cur = p.tile_act + (long long)par * p.tile_buf;
    uint8_t* const flags_nxt = p.tile_act + (long long)(par ^ 1) * p.tile_buf;
    const unsigned long long n_warps = (unsigned long long)gridDim.x * TILES_WARPS;
    CellT* const state = reinterpret_cast<CellT*>(p.state);
    constexpr int R = NSRC + 1;
    const int H = p.H, TX = p.tiles_x;

    auto dflush = [&]() {
        if (dcount == 0) return;
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.qcount + par, (unsigned long long)dcount);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < dcount; i += 32) {
            const unsigned long long slot = base + i;
            if (slot < (unsigned long long)p.qcap) p.queue[slot] = dq[i];
            else p.overflow[par] = 1;
        }
        dcount = 0;
        __syncwarp();
    };
    // change-log entries are staged per warp and appended in runs: the log lives in mapped host memory, and a run
    // of entries is a few full PCIe writes where single entries would be one small write each
    auto lflush = [&]() {
        if (lcount == 0) return;
        __syncwarp();
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(p.chg_count, (unsigned long long)lcount);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < lcount; i += 32) {
            const unsigned long long slot = base + i;
            if (slot < (unsigned long long)p.chg_cap) p.chg[slot] = lq[i];
            else *p.chg_overflow = 1;
        }
        lcount = 0;
        __syncwarp();
    };
    auto lpush = [&](bool have, long long idx, int burn_status) {
        const uint32_t m = __ballot_sync(0xffffffffu, have);
        if (!m) return;
        if (have) lq[lcount + __popc(m & lt)] = (unsigned long long)(idx + p.idx_base) | ((unsigned long long)burn_status << 48);
        lcount += __popc(m);
        if (lcount > WQ_CAP - 32) lflush();
    };
    // every lane calls; lanes with `have` contribute one item for k_eval
    auto dpush = [&](bool have, unsigned long long item) {
        const uint32_t m = __ballot_sync(0xffffffffu, have);
        if (!m) return;
        if (have) dq[dcount + __popc(m & lt)] = item;
        dcount += __popc(m);
        if (dcount > WQ_CAP - 32) dflush();
    };

    // appends to the next step's list: lane k < 9 may have set the "listed" byte of one neighbour tile with an
    // atomic whose answer (was it me who set it?) is only consumed here, a tile later
    // (all that stays in registers across the tile is the atomic's answer: ~0 = nothing pending / listed already;
    // the task and the byte's position in its word wait in shared memory)
    uint32_t pend_old = ~0u;
    unsigned long long* const pq = pq_all[warp];
    auto resolve = [&]() {
        const unsigned long long pend_task = pq[lane] & ~(3ull << 60);
        const bool won = ((pend_old >> (8 * (int)((pq[lane] >> 60) & 3))) & 1u) == 0;
        pend_old = ~0u;
        const uint32_t wm = __ballot_sync(0xffffffffu, won);
        if (!wm) return;
        unsigned int slot = 0;
        if (lane == 0) slot = atomicAdd(&aq_n, (unsigned int)__popc(wm));
        slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(wm & lt);
        if (won) {
            if (slot < AQ_CAP) {
                aq[slot] = pend_task;
            } else {  // the block's stage is full: straight to the list
                const unsigned long long g = atomicAdd(p.rows_count + (par ^ 1), 1ULL);
                if (g < (unsigned long long)p.rows_cap) tasks_nxt[g] = pend_task;
            }
        }
    };
    // tiles are dealt to the warps round-robin (a dynamic hand-out through a counter, and an L1 prefetch of the
    // warp's next tile, were measured and bought nothing: DESIGN.md section 4)
    for (unsigned long long ti = ti0; ti < n; ti += n_warps, task = tasks[min(ti, (unsigned long long)p.rows_cap - 1)]) {
        const int ty = (int)(task & 0xFFFFu), tx = (int)((task >> 16) & 0xFFFFu), env = (int)(task >> 32);
        EnvMeta* const mp = p.meta + (long long)par * p.meta_stride + env;
        const EnvMeta m = *mp;
        const long long tile_id = (long long)env * p.tile_stride + (long long)(ty * TX + tx);
        if (lane == 0) flags_cur[tile_id] = 0;  // off this step's list
        if (!m.running) continue;  // an env that has quit only comes back through a reset, which lists its tiles
        const int t = m.t;
        const bool spread = !m.time_quit;
        const int y0 = ty * 32, y = y0 + lane;
        const bool valid = y < H;
        const int x0 = tx * TW;  // column of bit 1
        const long long row0_idx = (long long)env * p.plane + (long long)y0 * p.pitch + x0 - 1;  // cell of (row 0, bit 0)
        // word (plane 0, row 0) of this word column; plane k is PL * k words further (PL < 2^28: sfb_create)
        uint32_t* const base = p.bits + (long long)env * p.bits_env + (long long)tx * H;
        const uint32_t PL = (uint32_t)p.bits_plane;
        // the two rows outside the tile ride in lanes 0 (row y0 - 1) and 1 (row y0 + 32)
        const int yh = lane == 0 ? y0 - 1 : y0 + 32;
        const bool hvalid = lane < 2 && yh >= 0 && yh < H;

        // ---- every word of the tile's window, all loads in flight together
        const int slot0 = (t - 1) % R;                      // plane of the sprites of duration 0 (t >= 1)
        const int e = slot0 + 1 == R ? 0 : slot0 + 1;       // = t mod R: duration max_fire_duration; this call's ignitions go here
        uint32_t c[NSRC], h[NSRC];
        {
            uint32_t off = (uint32_t)(BP_RING + slot0) * PL;
            const uint32_t wrap = (uint32_t)(R - 1) * PL;
            int slot = slot0;
#pragma unroll
            for (int a = 0; a < NSRC; ++a) {
                c[a] = h[a] = 0;
                if (valid) c[a] = base[off + (uint32_t)y];
                if (hvalid) h[a] = base[off + (uint32_t)yh];
                if (slot == 0) { slot = R - 1; off += wrap; } else { --slot; off -= PL; }
            }
        }
        // this lane's words of the IGN / LINE / expiring planes are base[o_ign], base[o_line], base[o_e]
        const uint32_t o_ign = (uint32_t)y, o_line = PL + (uint32_t)y, o_e = (uint32_t)(BP_RING + e) * PL + (uint32_t)y;
        uint32_t ign = 0, line = 0, ew = 0;
        if (valid) {
            ign = base[o_ign];
            line = base[o_line];
            ew = base[o_e] & OWN;
        }
        // lane k < 9 speaks for the tile at (ty + k / 3 - 1, tx + k % 3 - 1) when the next step's list is written:
        // its "already listed" byte is fetched with the planes (a stale 0 only costs the atomic)
        uint8_t nlisted = 1;  // (also for the lanes that speak for no tile)
        {
            const int ndy = lane / 3 - 1, ndx = lane % 3 - 1;
            const bool nvalid = lane < 9 && ty + ndy >= 0 && ty + ndy < p.tiles_y && tx + ndx >= 0 && tx + ndx < TX;
            if (nvalid) nlisted = *reinterpret_cast<volatile uint8_t*>(flags_nxt + (tile_id + (long long)(ndy * TX + ndx)));
        }
        resolve();  // the previous tile's appends
        nb[lane] = 0;
        if (STATS) {
            st_tiles += lane == 0;
            st_pruned += __popc(ew);
        }

        // ---- sprites that reached max_fire_duration: BURNED, out of the ring (fire.py:116-161)
        if (__any_sync(0xffffffffu, ew != 0)) {
            if (ew) {
                atomicAnd(base + o_e, ~ew);
                if ((ew & 2u) && tx > 0) atomicAnd(base + o_e - H, ~(1u << 31));
                if ((ew & (1u << TW)) && tx + 1 < TX) atomicAnd(base + o_e + H, ~1u);
                // a control line drawn over a burning cell (mitigation.py:77) burns out with its sprite
                if (ew & ign) base[o_ign] = (ign &= ~ew);
                if (ew & line) base[o_line] = (line &= ~ew);
                const long long ri = row0_idx + (long long)lane * p.pitch;
                for (uint32_t mm = ew; mm; mm &= mm - 1) state[ri + (__ffs(mm) - 1)] = (CellT)ST_BURNED;
            }
            if (p.track) {
                uint32_t mm = ew;
                while (__any_sync(0xffffffffu, mm != 0)) {
                    const bool have = mm != 0;
                    const int b = have ? __ffs(mm) - 1 : 0;
                    mm &= mm - 1;
                    lpush(have, row0_idx + (long long)lane * p.pitch + b, 2);  // BurnStatus.BURNED
                }
            }
        }

        // ---- candidate search: youngest sources first, then the reference's write order
        uint32_t und = spread ? ign : 0u, d0 = 0, d1 = 0, d2 = 0, live = 0, window = 0;
#pragma unroll
        for (int a = 0; a < NSRC; ++a) {
            const uint32_t ca = c[a];
            // no sprite of this duration anywhere in the window (tiles listed for a control line only, the far
            // side of a front), or nothing left to decide: next duration
            if (!__any_sync(0xffffffffu, (ca | h[a]) != 0)) continue;  // (lanes 0 / 1 carry the rows outside the tile)
            live |= ca & OWN;
            window |= ca | h[a];
            if (!__any_sync(0xffffffffu, und != 0)) continue;
            uint32_t cu = __shfl_up_sync(0xffffffffu, ca, 1), cd = __shfl_down_sync(0xffffffffu, ca, 1);
            const uint32_t hu = __shfl_sync(0xffffffffu, h[a], 0), hd = __shfl_sync(0xffffffffu, h[a], 1);
            if (lane == 0) cu = hu;
            if (lane == 31) cd = hd;
            auto take = [&](uint32_t src, int dir) {
                const uint32_t w = src & und;
                und &= ~w;
                if (dir & 1) d0 |= w;
                if (dir & 2) d1 |= w;
                if (dir & 4) d2 |= w;
            };
            // destination bit b <- source at column b + 1: shift right
            if (p.diagonal) take(cd >> 1, 5);
            take(cd, 6);
            if (p.diagonal) take(cd << 1, 7);
            take(ca >> 1, 4);
            take(ca << 1, 0);
            if (p.diagonal) take(cu >> 1, 3);
            take(cu, 2);
            if (p.diagonal) take(cu << 1, 1);
        }
        const bool has_window = __any_sync(0xffffffffu, window != 0);  // a sprite bit anywhere in the 34 x 32 window
        uint32_t cand = spread ? (ign & ~und) : 0u;
        if (STATS) st_cand += __popc(cand);
        if (__any_sync(0xffffffffu, live != 0) && lane == 0) mp->any_live = 1;  // fire.py:637
        const bool any_cand = __any_sync(0xffffffffu, cand != 0);
        if (any_cand && lane == 0) mp->any_cand = 1;  // fire.py:651

        // control lines no fire touches are attenuated like all others if the env gets past the early
        // return (fire.py:271-278, :651-652): a whole-env fact -> k_eval
        if (p.attenuate && spread) {
            uint32_t mm = line & ~cand;
            if (STATS) st_def += __popc(mm);
            while (__any_sync(0xffffffffu, mm != 0)) {
                const bool have = mm != 0;
                const int b = have ? __ffs(mm) - 1 : 0;
                mm &= mm - 1;
                const long long idx = row0_idx + (long long)lane * p.pitch + b;
                dpush(have, have ? make_item(idx, DIR_NONE, (int)state[idx] & 7) : 0ull);
            }
        }
        // a candidate that still carries a live sprite (a burning cell that was made ignitable again from
        // outside): if it re-ignites, its old ring bit has to go -- after every tile has read it
        if (__any_sync(0xffffffffu, (cand & live) != 0)) {
            uint32_t mm = cand & live;
            while (__any_sync(0xffffffffu, mm != 0)) {
                const bool have = mm != 0;
                const int b = have ? __ffs(mm) - 1 : 0;
                mm &= mm - 1;
                int slot = slot0, found = 0;
#pragma unroll
                for (int a = 0; a < NSRC; ++a) {
                    if ((c[a] >> b) & 1u) found = slot;
                    slot = slot == 0 ? R - 1 : slot - 1;
                }
                dpush(have, make_item(row0_idx + (long long)lane * p.pitch + b, DIR_UNRING, found));
            }
        }

        // ---- the tile's candidates, one per lane: rate look-up, burn +=, ignition (process_item)
        bool ignited_any = false;
        while (__any_sync(0xffffffffu, cand != 0)) {
            const int cnt = __popc(cand);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            const int total = min(__shfl_sync(0xffffffffu, incl, 31), WQ_CAP);
            int at = incl - cnt;
            while (cand && at < WQ_CAP) {  // the first WQ_CAP candidates in (row, column) order
                const int b = __ffs(cand) - 1;
                cand &= cand - 1;
                cq[at++] = (uint16_t)(lane | (b << 5));
            }
            __syncwarp();
            for (int i0 = 0; i0 < total; i0 += 32) {
                const bool have = i0 + lane < total;
                const int it = have ? cq[i0 + lane] : 0;
                const int row = it & 31, b = it >> 5;
                // the winning direction and the control-line bit sit in the registers of lane `row`
                const uint32_t r0 = __shfl_sync(0xffffffffu, d0, row), r1 = __shfl_sync(0xffffffffu, d1, row);
                const uint32_t r2 = __shfl_sync(0xffffffffu, d2, row), rl = __shfl_sync(0xffffffffu, line, row);
                bool ignited = false;
                long long idx = 0;
                if (have) {
                    idx = row0_idx + (long long)(row * p.pitch + b);
                    const int dir = (int)(((r0 >> b) & 1u) | (((r1 >> b) & 1u) << 1) | (((r2 >> b) & 1u) << 2));
                    // an ignitable cell that is no control line is UNBURNED; a line's kind is in its byte
                    const int st = ((rl >> b) & 1u) ? ((int)state[idx] & 7) : ST_UNBURNED;
                    ignited = process_item<CellT>(p, m, env, idx, dir, st);
                    if (ignited) atomicOr(&nb[row], 1u << b);
                }
                if (p.track) lpush(ignited, idx, 1);  // BurnStatus.BURNING
                ignited_any |= __any_sync(0xffffffffu, ignited);
            }
            __syncwarp();
        }

        // ---- this call's sprites join plane t mod R; they are no longer ignitable
        uint32_t mine = 0, all = 0, top = 0, bot = 0;
        if (ignited_any) {
            mine = nb[lane];
            if (STATS) st_ign += __popc(mine);
            if (mine) {
                atomicOr(base + o_e, mine);
                if ((mine & 2u) && tx > 0) atomicOr(base + o_e - H, 1u << 31);
                if ((mine & (1u << TW)) && tx + 1 < TX) atomicOr(base + o_e + H, 1u);
                base[o_ign] = ign & ~mine;
                if (line & mine) base[o_line] = (line &= ~mine);
            }
            all = __reduce_or_sync(0xffffffffu, mine);
            top = __shfl_sync(0xffffffffu, mine, 0);
            bot = __shfl_sync(0xffffffffu, mine, min(31, H - 1 - y0));
        }
        // ---- the next step's list: the tile itself if it still has something to look at, and the neighbours
        // whose window holds a cell that ignited
        const bool stay = has_window || __any_sync(0xffffffffu, (mine | (p.attenuate ? line : 0u)) != 0);
        {
            const int ndy = lane / 3 - 1, ndx = lane % 3 - 1;
            const uint32_t rowbits = ndy < 0 ? top : (ndy > 0 ? bot : all);          // what ignited next to that row of tiles
            const uint32_t colmask = ndx < 0 ? 2u : (ndx > 0 ? (1u << TW) : OWN);    // ... and next to that column
            bool want = (rowbits & colmask) != 0;
            if (lane == 4) want = stay;
            // the atomic is issued now and looked at while the next tile's loads are in flight (resolve)
            if (want && !nlisted) {  // (nlisted is 1 for tiles outside the grid and for lanes >= 9)
                const long long ntile = (long long)env * p.tile_stride + (long long)((ty + ndy) * TX + tx + ndx);
                pend_old = atomicOr(reinterpret_cast<uint32_t*>(flags_nxt) + (ntile >> 2), 1u << (8 * (int)(ntile & 3)));
                pq[lane] = make_tile_task(env, ty + ndy, tx + ndx) | ((unsigned long long)(ntile & 3) << 60);
            }
        }
    }
    resolve();
    dflush();
    lflush();
    // the block's appends, one atomic for all of them -- and, in a step without k_eval, the block's ticket (the
    // list entries themselves are only read by the next kernel: they need not be out before it)
    __shared__ unsigned int is_last;
    if (fuse) __threadfence();  // this block's env flags and queue items before its ticket
    __syncthreads();
    const unsigned int n_app = min(aq_n, (unsigned int)AQ_CAP);
    if (threadIdx.x == 0) {
        if (n_app) aq_base = (unsigned int)atomicAdd(p.rows_count + (par ^ 1), (unsigned long long)n_app);
        is_last = (fuse && atomicAdd(p.ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < n_app; i += blockDim.x)
        if ((long long)aq_base + i < p.rows_cap) tasks_nxt[aq_base + i] = aq[i];
    if (fuse) {
        if (is_last) {
            __threadfence();
            const long long nq = (long long)min(*reinterpret_cast<volatile unsigned long long*>(p.qcount + par), (unsigned long long)p.qcap);
            if (*reinterpret_cast<volatile int32_t*>(p.overflow + par)) {  // (a queue too short for the control lines: the dense form)
                const long long total = (long long)p.E * p.plane;
                for (long long i = threadIdx.x; i < total; i += blockDim.x) dense_cell<CellT>(p, par, i);
                if (p.track && threadIdx.x == 0) *p.chg_overflow = 1;
            } else {
                for (long long i = threadIdx.x; i < nq; i += blockDim.x) {
                    const unsigned long long it = *reinterpret_cast<volatile unsigned long long*>(p.queue + i);
                    const long long idx = (long long)(it & 0xFFFFFFFFFFFFull);
                    const int dir = (int)((it >> 48) & 0xF), st = (int)((it >> 52) & 7), env = (int)(idx / p.plane);
                    EnvMeta m = p.meta[(long long)par * p.meta_stride + env];
                    m.any_cand = reinterpret_cast<volatile EnvMeta*>(p.meta + (long long)par * p.meta_stride + env)->any_cand;
                    if (dir == DIR_UNRING) bits_unring<CellT>(p, m, env, idx, st);
                    else process_item<CellT>(p, m, env, idx, dir, st);  // DIR_NONE: never ignites, nothing to log
                }
            }
            __syncthreads();
            for (long long env = threadIdx.x; env < p.E; env += blockDim.x) advance_env(p, par, env);
            if (threadIdx.x == 0) {
                close_step(p, par);
                *p.ticket = 0;
            }
        }
    }
    if (STATS && p.tile_stats) {  // one atomic per warp and counter that has something
        const unsigned int v[5] = {st_tiles * (32u * TW), st_cand, st_ign, st_pruned, st_def};
        const int slot[5] = {0, 1, 2, 3, 6};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const unsigned int w = __reduce_add_sync(0xffffffffu, v[k]);
            if (lane == 0 && w) atomicAdd(p.tile_stats + slot[k], (unsigned long long)w);
        }
        const unsigned int w = __reduce_add_sync(0xffffffffu, st_tiles);
        if (lane == 0 && w) atomicAdd(p.tile_stats + 5, (unsigned long long)w);
    }
}

}  // namespace sfb
