// Bitboard front end (SFB_FRONT_BITS): the candidate search of one update() call as word-wide bit
// operations on 32 x 32-cell tiles, instead of a cell-by-cell look at 512-cell rows (k_rows).
//
// Next to the canonical per-cell state bytes the handle keeps bit planes, one 32-bit word per
// (env, tile column of 32 cells, row):
//
//   IGN    the cell is ignitable (UNBURNED or a control line, fire.py:192-205)
//   LINE   the cell is a control line
//   RING   R = max_fire_duration + 1 planes: plane (s mod R) holds the Fire sprites created by update()
//          call s.  In call t the sprites of duration a (fire.py:633) are plane (t - 1 - a) mod R: the
//          sources of the step are R - 1 of the planes as they stand, the remaining plane
//          (t mod R) holds the sprites that have reached max_fire_duration -- they are pruned
//          (fire.py:116-161), the plane is cleared, and the cells ignited by this call are written into
//          it.  No sprite code is ever decoded and no plane that is read as a source is written during
//          the step, so tiles never race.
//
// k_tiles, one warp per flagged tile, lane = row: for every duration (youngest first) and every
// neighbour in the reference's last-write-wins order (S-E, S, S-W, E, W, N-E, N, N-W: fire.py:704-705
// + sprite-list order) the shifted source plane is ANDed with the still undecided ignitable cells; three
// more planes collect the winning direction.  ~60 warp-instructions per duration decide all 1024 cells
// of a tile.  Candidates and (with attenuation) untouched control-line cells are pushed to the work
// queue exactly as k_rows pushes them, so k_eval is unchanged -- it only also sets the new sprite's ring
// bit and clears its IGN / LINE bits when a cell ignites.
//
// A tile is flagged (tile_act) while its 34 x 34 window holds a ring bit or (with attenuation) the tile
// holds a control line: raised at ignitions / resets / mitigation / map uploads, lowered by k_tiles.
#pragma once
#include "sfb_kernels.cuh"

namespace sfb {

constexpr int BP_IGN = 0, BP_LINE = 1, BP_RING = 2;
constexpr int BITS_MAX_DUR = 7;  // ring of at most 8 planes; longer-lived sprites use the byte front ends

__device__ __forceinline__ uint32_t* bits_word(const DevParams& p, int env, int plane, int tx, int y) {
    return p.bits + (long long)env * p.bits_env + (long long)plane * p.bits_plane + (long long)tx * p.H + y;
}
// plane of the sprites that have duration `a` in update() call t (1-based)
__device__ __forceinline__ int ring_slot(int t, int a, int R) {
    int s = (t - 1 - a) % R;
    return s < 0 ? s + R : s;
}
__device__ __forceinline__ void mark_tiles_around(const DevParams& p, int env, int y, int x) {
    const int ty0 = (y > 0 ? y - 1 : 0) >> 5, ty1 = (y + 1 < p.H ? y + 1 : p.H - 1) >> 5;
    const int tx0 = (x > 0 ? x - 1 : 0) >> 5, tx1 = (x + 1 < p.W ? x + 1 : p.W - 1) >> 5;
    uint8_t* f = p.tile_act + (long long)env * p.tile_stride;
    for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) f[ty * p.tiles_x + tx] = 1;
}
// fire_map[y, x] was set to internal status `s` from outside (mitigation.py:77): sprite planes untouched
__device__ __forceinline__ void bits_on_status(const DevParams& p, int env, int y, int x, int s) {
    const uint32_t bit = 1u << (x & 31);
    uint32_t* ign = bits_word(p, env, BP_IGN, x >> 5, y);
    uint32_t* line = bits_word(p, env, BP_LINE, x >> 5, y);
    if (ignitable(s)) atomicOr(ign, bit);
    else atomicAnd(ign, ~bit);
    if (s & ST_LINE_BIT) {
        atomicOr(line, bit);
        if (p.attenuate) p.tile_act[(long long)env * p.tile_stride + (y >> 5) * p.tiles_x + (x >> 5)] = 1;
    } else {
        atomicAnd(line, ~bit);
    }
}

// (re)derive every plane of n envs (a device list, or [env0, env0 + n)) from their state bytes; one
// thread per word.  Durations are those the NEXT update() call will see.
template <typename CellT>
__global__ void k_bits_rebuild(const DevParams p, const int par, const int32_t* envs, const int env0, const int n) {
    const long long per_env = (long long)p.tiles_x * p.H, total = (long long)n * per_env;
    const CellT* state = reinterpret_cast<const CellT*>(p.state);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / per_env);
        const int env = envs ? envs[k] : env0 + k;
        const long long r = i - (long long)k * per_env;
        const int y = (int)(r % p.H), tx = (int)(r / p.H);  // consecutive threads: consecutive rows of one tile column
        const int t = p.meta[(long long)par * p.meta_stride + env].t;
        const int tm1 = (t - 1) % Cell<CellT>::M;
        const CellT* row = state + (long long)env * p.plane + (long long)y * p.pitch + tx * 32;
        uint32_t ign = 0, line = 0, ring[BITS_MAX_DUR + 1];
        for (int s = 0; s <= BITS_MAX_DUR; ++s) ring[s] = 0;
        const int nx = min(32, p.W - tx * 32);
        for (int b = 0; b < nx; ++b) {
            const int c = row[b];
            const int st = c & 7;
            if (ignitable(st)) ign |= 1u << b;
            if (st & ST_LINE_BIT) line |= 1u << b;
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
        if (any) {  // flag every tile whose 34 x 34 window can see this word (generously)
            mark_tiles_around(p, env, y, tx * 32);
            mark_tiles_around(p, env, y, min(tx * 32 + 31, p.W - 1));
        }
    }
}

// tile task: ty | tx << 16 | env << 32
__device__ __forceinline__ unsigned long long make_tile_task(int env, int ty, int tx) {
    return (unsigned long long)(unsigned)ty | ((unsigned long long)(unsigned)tx << 16) | ((unsigned long long)(unsigned)env << 32);
}

// compacts the flagged tiles of running envs into this step's task list (the same scan as k_row_list)
__global__ void __launch_bounds__(256) k_tile_list(const DevParams p, const int par) {
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n4 = (long long)p.E * p.tile_stride / 4;  // tile_stride is a multiple of 4
    const long long used = (long long)p.tiles_y * p.tiles_x;
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(p.tile_act);
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; base < n4; base += stride) {
        const long long w = base + lane;
        const uint32_t f = w < n4 ? flags[w] : 0u;
        if (!__any_sync(0xffffffffu, f != 0)) continue;
        unsigned long long task[4];
        int cnt = 0;
        if (f) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (!((f >> (8 * b)) & 0xFFu)) continue;
                const long long u = 4 * w + b;
                const int env = (int)(u / p.tile_stride);
                if (!p.meta[(long long)par * p.meta_stride + env].running) continue;
                const long long r = u - (long long)env * p.tile_stride;
                if (r >= used) continue;  // pad flag (a map upload sets whole envs, pads included)
                const int ty = (int)(r / p.tiles_x);
                task[cnt++] = make_tile_task(env, ty, (int)(r - (long long)ty * p.tiles_x));
            }
        }
        int incl = cnt;
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

constexpr int TILES_WARPS = 4;

template <typename CellT>
__global__ void __launch_bounds__(TILES_WARPS * 32) k_tiles(const DevParams p, const int par) {
    using C = Cell<CellT>;
    __shared__ unsigned long long wq_all[TILES_WARPS][WQ_CAP];  // per-warp staging of work items
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    unsigned long long* const wq = wq_all[warp];
    int wcount = 0;
    const unsigned long long n = min(p.rows_count[par], (unsigned long long)p.rows_cap);
    const unsigned long long n_warps = (unsigned long long)gridDim.x * TILES_WARPS;
    CellT* const state = reinterpret_cast<CellT*>(p.state);
    const int R = p.ring, H = p.H, TX = p.tiles_x;

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
    // every lane pushes the cells of `mask` (bits of its row) as work items, a bit per round
    auto push_bits = [&](uint32_t mask, uint32_t d0, uint32_t d1, uint32_t d2, uint32_t line, bool with_dir, int fixed_dir,
                         long long row_idx, int x0) {
        while (__any_sync(0xffffffffu, mask != 0)) {
            const bool have = mask != 0;
            const int b = have ? __ffs(mask) - 1 : 0;
            mask &= mask - 1;
            const uint32_t m = __ballot_sync(0xffffffffu, have);
            if (have) {
                const long long idx = row_idx + x0 + b;
                const int dir = with_dir ? (int)(((d0 >> b) & 1u) | (((d1 >> b) & 1u) << 1) | (((d2 >> b) & 1u) << 2)) : fixed_dir;
                // an ignitable cell that is no control line is UNBURNED; a line's kind is in its byte
                const int s = ((line >> b) & 1u) ? ((int)state[idx] & 7) : (fixed_dir == DIR_PRUNED ? ST_BURNED : ST_UNBURNED);
                wq[wcount + __popc(m & lt)] = make_item(idx, dir, s);
            }
            wcount += __popc(m);
            if (wcount > WQ_CAP - 32) flush();
        }
    };

    for (unsigned long long ti = (unsigned long long)blockIdx.x * TILES_WARPS + warp; ti < n; ti += n_warps) {
        const unsigned long long task = p.rows[ti];
        const int ty = (int)(task & 0xFFFFu), tx = (int)((task >> 16) & 0xFFFFu), env = (int)(task >> 32);
        EnvMeta* const mp = p.meta + (long long)par * p.meta_stride + env;
        const int t = mp->t;
        const bool spread = !mp->time_quit;
        const int y = ty * 32 + lane;
        const bool valid = y < H;
        const int x0 = tx * 32;
        const long long row_idx = (long long)env * p.plane + (long long)y * p.pitch;  // cell index of (y, x = 0)
        const uint32_t* const base = p.bits + (long long)env * p.bits_env;
        const bool has_l = tx > 0, has_r = tx + 1 < TX;
        auto word = [&](int plane, int txx, int yy) -> uint32_t { return base[(long long)plane * p.bits_plane + (long long)txx * H + yy]; };

        uint32_t ign = valid ? word(BP_IGN, tx, y) : 0u;
        uint32_t line = valid ? word(BP_LINE, tx, y) : 0u;

        // sprites that reached max_fire_duration: BURNED, out of the ring (fire.py:116-161)
        const int e = ring_slot(t, R - 1, R);  // = t mod R: duration max_fire_duration; this call's ignitions go here
        uint32_t ew = valid ? word(BP_RING + e, tx, y) : 0u;
        uint32_t window = ew;  // any ring bit in the 34 x 34 window (for lowering the flag)
        uint32_t pruned = 0;
        if (__any_sync(0xffffffffu, ew != 0)) {
            for (uint32_t m = ew; m; m &= m - 1) {
                const int b = __ffs(m) - 1;
                const long long idx = row_idx + x0 + b;
                const int c = state[idx];
                // a bit whose sprite was replaced by a newer one on the same cell (a line drawn over a burning
                // cell that re-ignited) is stale: the byte carries the newer sprite's code
                if ((c >> 3) != 0 && sprite_age<CellT>(c >> 3, (t - 1) % C::M) >= p.max_dur) {
                    state[idx] = (CellT)ST_BURNED;
                    pruned |= 1u << b;
                }
            }
            if (ew != 0) *bits_word(p, env, BP_RING + e, tx, y) = 0u;
            // a control line drawn over a burning cell (mitigation.py:77) burns out with its sprite (fire.py:157-159)
            if (pruned & ign) *bits_word(p, env, BP_IGN, tx, y) = (ign &= ~pruned);
            if (pruned & line) *bits_word(p, env, BP_LINE, tx, y) = (line &= ~pruned);
            if (p.track) push_bits(pruned, 0, 0, 0, 0, false, DIR_PRUNED, row_idx, x0);
        }

        // candidate search: youngest sources first, then the reference's write order
        uint32_t und = spread ? ign : 0u, d0 = 0, d1 = 0, d2 = 0, live = 0;
        for (int a = 0; a < R - 1; ++a) {
            const int k = BP_RING + ring_slot(t, a, R);
            uint32_t c = 0, lr = 0;  // own word; bit 0: the cell left of the tile, bit 1: the cell right of it
            if (valid) {
                c = word(k, tx, y);
                if (has_l) lr |= word(k, tx - 1, y) >> 31;
                if (has_r) lr |= (word(k, tx + 1, y) & 1u) << 1;
            }
            uint32_t cu = __shfl_up_sync(0xffffffffu, c, 1), lru = __shfl_up_sync(0xffffffffu, lr, 1);
            uint32_t cd = __shfl_down_sync(0xffffffffu, c, 1), lrd = __shfl_down_sync(0xffffffffu, lr, 1);
            if (lane == 0) {  // the row above the tile
                cu = lru = 0;
                if (y > 0 && valid) {
                    cu = word(k, tx, y - 1);
                    if (has_l) lru |= word(k, tx - 1, y - 1) >> 31;
                    if (has_r) lru |= (word(k, tx + 1, y - 1) & 1u) << 1;
                }
            }
            if (lane == 31) {  // the row below it
                cd = lrd = 0;
                if (y + 1 < H) {
                    cd = word(k, tx, y + 1);
                    if (has_l) lrd |= word(k, tx - 1, y + 1) >> 31;
                    if (has_r) lrd |= (word(k, tx + 1, y + 1) & 1u) << 1;
                }
            }
            live |= c;
            window |= c | lr | cu | lru | cd | lrd;
            if (!__any_sync(0xffffffffu, und != 0)) continue;
            // destination bit x <- source at (x + 1): shift right, the tile to the right supplies bit 31
            const uint32_t sE = (c >> 1) | ((lr >> 1) << 31), sW = (c << 1) | (lr & 1u);
            const uint32_t sSE = (cd >> 1) | ((lrd >> 1) << 31), sSW = (cd << 1) | (lrd & 1u);
            const uint32_t sNE = (cu >> 1) | ((lru >> 1) << 31), sNW = (cu << 1) | (lru & 1u);
            auto take = [&](uint32_t src, int dir) {
                const uint32_t w = src & und;
                und &= ~w;
                if (dir & 1) d0 |= w;
                if (dir & 2) d1 |= w;
                if (dir & 4) d2 |= w;
            };
            if (p.diagonal) take(sSE, 5);
            take(cd, 6);
            if (p.diagonal) take(sSW, 7);
            take(sE, 4);
            take(sW, 0);
            if (p.diagonal) take(sNE, 3);
            take(cu, 2);
            if (p.diagonal) take(sNW, 1);
        }
        const uint32_t cand = spread ? (ign & ~und) : 0u;
        if (__any_sync(0xffffffffu, live != 0) && lane == 0) mp->any_live = 1;  // fire.py:637
        if (__any_sync(0xffffffffu, cand != 0)) {
            if (lane == 0) mp->any_cand = 1;  // fire.py:651
            push_bits(cand, d0, d1, d2, line, true, 0, row_idx, x0);
        }
        // control lines no fire touches are attenuated like all others if the env gets past the early
        // return (fire.py:271-278, :651-652): deferred items, k_eval decides
        const bool att_lines = p.attenuate && spread;
        if (att_lines && __any_sync(0xffffffffu, (line & ~cand) != 0)) push_bits(line & ~cand, 0, 0, 0, line, false, DIR_NONE, row_idx, x0);
        // nothing left to look at: the tile leaves the list until an ignition / line / upload flags it again
        if (!__any_sync(0xffffffffu, (window | (p.attenuate ? line : 0u)) != 0) && lane == 0)
            p.tile_act[(long long)env * p.tile_stride + (long long)ty * TX + tx] = 0;
    }
    flush();
}

}  // namespace sfb
