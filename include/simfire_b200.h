/*
 * simfire_b200.h -- C ABI of the B200 fire-spread stepper (libsimfire_b200.so).
 *
 * The reference (mitrefireline/simfire, pure Python) has no FFI: the seam this library
 * replaces is the manager object that FireSimulation constructs and steps,
 *     simfire/sim/simulation.py:273-291   _create_fire()  -> RothermelFireManager(...)
 *     simfire/sim/simulation.py:533-535   run(): fire_map, status = fire_manager.update(fire_map)
 *     simfire/game/managers/fire.py:293   RothermelFireManager.__init__
 *     simfire/game/managers/fire.py:616   RothermelFireManager.update
 * Every entry point below names the reference interface it stands in for.  All pointers
 * are plain host pointers unless a comment says "device"; buffers are borrowed for the
 * duration of the call only.  No CUDA, C++ or torch types cross this boundary.
 *
 * One handle = one CUDA device = E independent simulations ("envs") of one H x W grid,
 * or (slab mode, sfb_params.slab_*) one horizontal slab of a larger grid whose other
 * slabs live in other handles / processes / GPUs.
 *
 * Error convention: every function returns 0 on success or a negative sfb_error; the
 * message of the last failure on the calling thread is sfb_last_error().  Nothing throws.
 * Threading: calls on one handle must be serialised by the caller (the reference is
 * single-threaded); different handles are independent.
 *
 * Environment variables read by sfb_create / sfb_sync_fire_maps (measurement and test knobs; none of
 * them changes a result): SFB_UNIT_SKIP=0|1, SFB_UNIT_ROWS=0|1, SFB_GROUP_GRAPH=0|1 override the
 * corresponding sfb_flags; SFB_SWEEP_BLOCKS_PER_SM=n; SFB_HOST_THREADS=n (threads that patch the host
 * mirror, default: all cores); SFB_PATCH_PARALLEL_MIN=n (shortest log patched by several threads);
 * SFB_DEBUG_TIMING=1 (sfb_sync_fire_maps prints where its time went).
 */
#ifndef SIMFIRE_B200_H
#define SIMFIRE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_ABI_VERSION 1

typedef struct sfb_sim sfb_sim; /* opaque handle */

enum sfb_error {
    SFB_OK = 0,
    SFB_ERR_INVALID = -1, /* bad argument */
    SFB_ERR_CUDA = -2,    /* CUDA runtime error (message has the CUDA string) */
    SFB_ERR_NOMEM = -3,   /* device or host allocation failed */
    SFB_ERR_STATE = -4    /* call not valid in the handle's current state */
};

/* BurnStatus values as they appear in fire_map (simfire/enums.py:52-69). */
enum sfb_burn_status {
    SFB_UNBURNED = 0,
    SFB_BURNING = 1,
    SFB_BURNED = 2,
    SFB_FIRELINE = 3,
    SFB_SCRATCHLINE = 4,
    SFB_WETLINE = 5
};

/* GameStatus values returned by update() (simfire/enums.py:106-115). */
enum sfb_game_status { SFB_QUIT = 0, SFB_RUNNING = 1 };

/* sfb_params.flags */
enum sfb_flags {
    SFB_DIAGONAL_SPREAD = 1,   /* fire.py:211-221 (8 neighbours) vs :223-228 (4)            */
    SFB_ATTENUATE_LINE_ROS = 2,/* fire.py:271-278 vs :280-282                                */
    SFB_SHARED_STATIC = 4,     /* all envs read one set of static planes (one terrain)       */
    SFB_KEEP_ROS = 8,          /* materialise the dense rate_of_spread plane (fire.py:704-708);
                                  costs a dense 8 B/cell pass per step -- parity tests only  */
    SFB_HAS_MAX_TIME = 16,     /* max_time is not None (fire.py:641)                         */
    SFB_WIDE_CELLS = 32,       /* use the 16-bit cell layout even when max_fire_duration <= 30
                                  (it is selected automatically above that); tests only      */
    SFB_SWEEP_LDG = 64,        /* stream the state with 128-bit global loads instead of the TMA
                                  ring (A/B measurements; slab mode always uses it)          */
    SFB_TRACK_CHANGES = 128,   /* keep a device log of every BurnStatus change so that
                                  sfb_sync_fire_maps can patch a host mirror incrementally   */
    SFB_KEEP_IGNITION = 256    /* keep an int32 plane with the update() call that ignited each
                                  cell: enough to rebuild the fire-spread graph the reference
                                  maintains per step (graph.py:84-150, called at fire.py:584)   */,
    SFB_UNIT_SKIP_OFF = 512,   /* always sweep every (env, rows, columns) unit                      */
    SFB_UNIT_SKIP_ON = 1024,   /* keep per-unit activity flags and look only at the flagged units even
                                  for small handles (default: on from 1024 units up, off in slab
                                  mode).  Results are identical either way; the reference has no
                                  counterpart (it walks its sprite list, fire.py:655-690)          */
    SFB_UNIT_CHUNKS = 2048,    /* with unit skipping: a unit is a chunk of rows of one strip and the
                                  flagged units are swept (k_units + k_sweep).  Default: a unit is a
                                  single row of a strip, the flagged units are the row tasks
                                  themselves and nothing is swept (k_row_list)                     */
    SFB_FRONT_LISTS = 8192,    /* the list-driven step (sfb_lists.cuh): one watch list of the cells that can
                                  change (sprites, their ignitable neighbours, attenuated control lines), one
                                  kernel per step (k_front), no env groups.  Fewer instructions and bytes per
                                  step than the sweep front ends, but all of it scattered 32-byte reads: it pays
                                  where the planes fit L2 (single envs, small batches).  Results are identical  */
    SFB_FRONT_BITS = 16384,    /* the bitboard step (sfb_bits.cuh): bit planes next to the state bytes (ignitable,
                                  control line, one plane per sprite duration), the candidate search as word-wide
                                  bit operations on tiles of 32 rows x 30 columns, candidates evaluated by the
                                  same warp, the next step's tile list written by the step itself: k_tiles +
                                  k_eval, all envs as one group.  Needs max_fire_duration <= 7; results are
                                  identical.  It is what a handle of >= 1024 row units gets when none of the
                                  front-end flags (SFB_UNIT_*, SFB_SWEEP_LDG, SFB_FRONT_LISTS, rows_per_chunk,
                                  slab mode) is given; the flag forces it on smaller handles                  */
    SFB_STEP_GRAPH = 4096      /* multi-group handles: replay pairs of steps of sfb_step(n) as one CUDA
                                  graph forked over the group streams instead of enqueueing every
                                  kernel (single-group handles always replay a graph).  Off by
                                  default: the join after every pair costs more overlap between the
                                  groups than the saved launches give back (measured)               */
};

/* The eight static per-cell inputs of the Rothermel evaluation, in the order of the
 * reference's parameter list (fire.py:481-497, rothermel.py:4-22).  Values are float32:
 * the reference casts them at fire.py:537 / :546 before any arithmetic. */
enum sfb_static_plane {
    SFB_W_0 = 0,
    SFB_DELTA = 1,
    SFB_M_X = 2,
    SFB_SIGMA = 3,
    SFB_U = 4,
    SFB_U_DIR = 5,
    SFB_SLOPE_MAG = 6,
    SFB_SLOPE_DIR = 7,
    SFB_N_STATIC = 8
};

/* Planes readable with sfb_get_plane (parity tests, RothermelFireManager attributes). */
enum sfb_state_plane {
    SFB_PLANE_BURN = 0, /* float64 burn_amounts (fire.py:370, :710)                     */
    SFB_PLANE_ROS = 1,  /* float64 rate_of_spread of the last step (needs SFB_KEEP_ROS) */
    SFB_PLANE_AGE = 2,  /* int32: -1 no sprite, else the sprite's duration (fire.py:633) */
    SFB_PLANE_STATUS = 3,/* int8 BurnStatus -- same as sfb_get_fire_map                 */
    SFB_PLANE_IGNITION = 4 /* int32: update() call that ignited the cell (0 = initial fire,
                              -1 = never ignited); needs SFB_KEEP_IGNITION                 */
};

/* Constructor arguments: RothermelFireManager.__init__ (fire.py:293-307) plus the
 * FuelParticle constants (simfire/world/parameters.py:8-27) and Environment.M_f (:53). */
typedef struct sfb_params {
    int32_t abi_version;       /* SFB_ABI_VERSION */
    int32_t device;            /* CUDA device ordinal */
    int32_t H, W;              /* screen_size = (H, W); arrays are indexed [y][x] */
    int32_t E;                 /* number of independent simulations in this handle */
    int32_t max_fire_duration; /* fire.py:116-161; 1..8189 */
    int32_t flags;             /* sfb_flags */
    int32_t rows_per_chunk;    /* 0 = auto; tuning knob of the sweep kernel */
    double pixel_scale;        /* ft per pixel, ignition threshold (fire.py:568) */
    double update_rate;        /* minutes per step (fire.py:696, :717) */
    double max_time;           /* minutes; used when SFB_HAS_MAX_TIME */
    float h, S_T, S_e, p_p;    /* FuelParticle */
    float M_f;                 /* Environment.M_f */
    int32_t env_groups;        /* 0 = auto; envs are stepped as this many independent groups on
                                  separate CUDA streams (k_sweep of one overlaps k_rows / k_eval of
                                  another); 1 in slab mode and for bitboard handles */
    int64_t queue_capacity;    /* 0 = auto; work-queue entries (8 B each) */
    /* Slab mode (single huge grid split in rows across handles): this handle holds rows
     * [slab_y0, slab_y0 + H) of a grid with slab_total_H rows.  0/0 = whole grid. */
    int32_t slab_y0, slab_total_H;
} sfb_params;

/* ---- lifetime -------------------------------------------------------------------- */

/* RothermelFireManager.__init__ (fire.py:293-380).  Allocates all device state.  Envs
 * start QUIT with an empty map until sfb_reset() places the initial fire. */
int sfb_create(const sfb_params* params, sfb_sim** out);
void sfb_destroy(sfb_sim* sim);
const char* sfb_last_error(void);
int sfb_abi_version(void);

/* ---- static inputs --------------------------------------------------------------- */

/* terrain.fuels / U / U_dir / slope planes (fire.py:367-378, :436-449).  `host` is a
 * float32 [H][W] array for one plane of env `env` (env = -1: every env, or the single
 * shared set under SFB_SHARED_STATIC). */
int sfb_set_static(sfb_sim* sim, int32_t env, int32_t plane, const float* host);
/* All eight planes at once: float32 [8][H][W] in sfb_static_plane order. */
int sfb_set_static_all(sfb_sim* sim, int32_t env, const float* host);

/* RothermelFireManager._compute_slopes (fire.py:436-449) on the device: from float64
 * elevations [H][W] (ft) computes np.gradient(elevations, pixel_scale) (central differences,
 * one-sided at the borders), slope_mag = sqrt(gx^2 + gy^2) and slope_dir = atan2(gy, gx + 1e-6)
 * in float64 and stores them as the SFB_SLOPE_MAG / SFB_SLOPE_DIR planes of env `env`
 * (-1: all / shared).  Not available in slab mode (the gradient crosses slab borders). */
int sfb_set_elevation(sfb_sim* sim, int32_t env, const double* elevations);

/* ---- between-step mutations ------------------------------------------------------ */

/* FireSimulation.reset() -> _create_fire_map + _create_fire (simulation.py:202-214,
 * :273-291, :555-566): map all UNBURNED, burn 0, one BURNING sprite at (x, y), elapsed 0,
 * status RUNNING.  `envs` lists n env indices (NULL = envs 0..n-1); xy = n pairs (x, y).
 * In slab mode y is a global row; slabs that do not contain it just clear. */
int sfb_reset(sfb_sim* sim, const int32_t* envs, int32_t n, const int32_t* xy);

/* ControlLineManager.update (mitigation.py:60-80) batched: points are (env, x, y, kind)
 * int32 quadruples, kind = BurnStatus value; fire_map[y, x] = kind unconditionally.
 * `points` is copied before the call returns; the writes themselves are enqueued on the
 * handle's stream, ahead of whatever is called next. */
int sfb_apply_points(sfb_sim* sim, const int32_t* points, int64_t n);

/* FireSimulation.load_mitigation (simulation.py:425-447) / the fire_map argument of
 * update(): replace the BurnStatus of every cell of envs [env0, env0+n) from int8
 * [n][H][W]; sprites (burning durations) are kept, as in the reference where the sprite
 * list is independent of fire_map (fire.py:101-103). */
int sfb_set_fire_map(sfb_sim* sim, int32_t env0, int32_t n, const int8_t* maps);

/* ---- the hot path ---------------------------------------------------------------- */

/* n_steps x RothermelFireManager.update (fire.py:616-719) on every RUNNING env.  Work is
 * enqueued on the handle's stream; returns after enqueueing unless sync != 0. */
int sfb_step(sfb_sim* sim, int32_t n_steps, int32_t sync);

/* Same, bracketed by CUDA events on the handle's stream; *ms = device time of the n steps. */
int sfb_step_timed(sfb_sim* sim, int32_t n_steps, float* ms);

/* Drop-in for `fire_map, status = manager.update(fire_map)` (simulation.py:535) with
 * HOST buffers: uploads int8 maps [n][H][W] of envs [env0, env0+n) (the caller may have
 * edited them, mitigation.py:77), runs one step, downloads the maps in place and writes
 * GameStatus per env to status[n] (may be NULL). */
int sfb_update(sfb_sim* sim, int32_t env0, int32_t n, int8_t* maps_inout, int32_t* status);

/* Drop-in for `fire_map = ConstantSpreadFireManager.update(fire_map)` (fire.py:754-787) with HOST buffers,
 * exactly as the reference executes it: sprites past max_fire_duration turn BURNED (fire.py:116-161), then
 * every sprite whose duration equals rate_of_spread sets its in-bounds ignitable neighbours (8, or 4 without
 * SFB_DIAGONAL_SPREAD) to BURNING.  The reference appends Fire sprites for those cells without a duration
 * entry, and its next prune slices them off again (fire.py:148-155): they stay BURNING and never spread;
 * here they get the status and no sprite.  Uploads int8 maps [n][H][W] of envs [env0, env0+n), advances EVERY
 * env of the handle by one call, downloads the maps in place.  No GameStatus (the reference returns none). */
int sfb_constant_spread_update(sfb_sim* sim, int32_t env0, int32_t n, int8_t* maps_inout, int32_t rate_of_spread);

int sfb_synchronize(sfb_sim* sim);

/* ---- results --------------------------------------------------------------------- */

/* fire_map of envs [env0, env0+n) as int8 BurnStatus [n][H][W]. */
int sfb_get_fire_map(sfb_sim* sim, int32_t env0, int32_t n, int8_t* out);
/* Bring a HOST mirror of every env's fire_map (int8 BurnStatus [E][H][W]) up to date.  The
 * first call, a call with a different buffer, or a call after sfb_set_fire_map / a log
 * overflow downloads everything (like sfb_get_fire_map); otherwise, with SFB_TRACK_CHANGES,
 * only the cells that changed since the previous call travel over PCIe and are patched into
 * `mirror` by host threads -- what `sim.fire_map` needs after each `run()` without copying
 * E*H*W bytes per step.  The caller must not modify `mirror` between calls.  *n_changes (may be
 * NULL) receives the number of patched cells, or -1 for a full download. */
int sfb_sync_fire_maps(sfb_sim* sim, int8_t* mirror, int64_t* n_changes);

/* Pause / resume the change log of a handle created with SFB_TRACK_CHANGES (a rollout that
 * only consumes the device-resident observation does not need it; logging costs PCIe writes).
 * Resuming forces the next sfb_sync_fire_maps to download everything once. */
int sfb_set_tracking(sfb_sim* sim, int32_t enabled);

/* One [H][W] plane of one env; element type per sfb_state_plane. */
int sfb_get_plane(sfb_sim* sim, int32_t env, int32_t plane, void* out);
/* Per-env GameStatus (int32), elapsed_time (float64, fire.py:717) and number of update()
 * calls made (int32); any pointer may be NULL. */
int sfb_get_status(sfb_sim* sim, int32_t* status, double* elapsed, int32_t* steps);

/* Zero-copy observation: after this call *dev points at DEVICE memory holding int8
 * BurnStatus [E][H][W], refreshed from the packed state on the handle's stream (the call
 * synchronises the stream before returning). */
int sfb_fire_map_device(sfb_sim* sim, void** dev);

/* Zero-copy view of the static inputs (FireSimulation.get_attribute_data, simulation.py:376-403): *records
 * points at DEVICE memory holding float32 records {w_0, delta, M_x, sigma, U, U_dir, slope_mag, slope_dir}
 * (sfb_static_plane order, 32 bytes per cell) laid out [n_sets][H][pitch_cells]; n_sets is 1 under
 * SFB_SHARED_STATIC, else E.  Plane k of set e is a strided view: element (y, x) at float index
 * ((e * plane_cells + y * pitch_cells + x) * 8 + k).  Read-only for the caller. */
int sfb_static_device(sfb_sim* sim, void** records, int64_t* plane_cells, int32_t* pitch_cells, int32_t* n_sets);

/* ---- slab mode: one large grid split in horizontal slabs across handles / GPUs ----------
 * Each handle is created with slab_y0 / slab_total_H and holds rows [slab_y0, slab_y0 + H).
 * The sweep kernel reads the row above / below its slab straight from the neighbour slab's
 * state plane (peer device memory over NVLink when the slabs live on different GPUs), so
 * there is no halo copy; what must be coordinated per step is
 *     (all slabs finished step t-1)  ->  sfb_step_sweep on every slab
 *     OR of the per-env flags across slabs (sfb_flags_device, an all-reduce MAX of int32)
 *     ->  sfb_step_eval on every slab.
 * simfire_b200/slab.py does this with torch.distributed (one process per GPU). */

/* Device pointer and geometry of the packed state plane [E][H][pitch] of this handle. */
int sfb_state_device(sfb_sim* sim, void** state, int64_t* plane_cells, int32_t* pitch_cells, int32_t* cell_bytes);
/* CUDA IPC handle (64 bytes) of the state plane, to be opened in the neighbours' processes. */
int sfb_ipc_export(sfb_sim* sim, void* handle64);
int sfb_ipc_open(int32_t device, const void* handle64, void** dev_ptr);
int sfb_ipc_close(int32_t device, void* dev_ptr);
/* top_row: device pointer to row (slab_y0 - 1) of env 0 inside the slab above (NULL: grid
 * edge); top_plane_cells: per-env stride of that slab's plane; same for the slab below. */
int sfb_set_halo(sfb_sim* sim, const void* top_row, int64_t top_plane_cells, const void* bottom_row,
                 int64_t bottom_plane_cells);
/* The two halves of sfb_step. */
int sfb_step_sweep(sfb_sim* sim);
int sfb_step_eval(sfb_sim* sim);
/* int32 view of the EnvMeta records of the step in flight (valid between sfb_step_sweep and
 * sfb_step_eval): 8 int32 per env; an element-wise MAX across slabs ORs any_live / any_cand
 * and leaves the other fields (identical on every slab) unchanged. */
int sfb_flags_device(sfb_sim* sim, void** flags, int64_t* n_int32);
/* Which half (0 / 1) of the double-buffered per-env records the NEXT step reads (the pointer
 * sfb_flags_device returns is that half; the other half starts n_int32 * 4 bytes before or after it). */
int sfb_get_parity(sfb_sim* sim, int32_t* parity);
/* Peer-memory coordination (no host or NCCL round trip per step).  The handle owns a small
 * mailbox (sfb_slab_mailbox: device pointer + byte offset from the state plane, so it can be
 * reached through the state plane's IPC mapping); sfb_slab_connect receives, for every slab q of
 * the grid, a device pointer to q's mailbox as seen from this device (own included).  After that
 * sfb_step_slab(n) enqueues n x [sweep, flag exchange, eval, done handshake] on the stream: the
 * handshakes are single-warp kernels that store into the peers' mailboxes and poll their own.
 * Every slab of the grid must call sfb_step_slab with the same n.  A wait that is not satisfied
 * within about a second sets an error that the next sfb_synchronize reports (SFB_ERR_STATE). */
int sfb_slab_mailbox(sfb_sim* sim, void** mailbox, int64_t* offset_from_state);
int sfb_slab_connect(sfb_sim* sim, int32_t rank, int32_t world, void* const* peer_mailboxes);
int sfb_step_slab(sfb_sim* sim, int32_t n_steps);

/* Run the handle's work on the caller's stream (cudaStream_t as void*), e.g. the stream a
 * communication library orders its collectives on.  NULL restores the handle's own stream. */
int sfb_set_stream(sfb_sim* sim, void* stream);

/* ---- introspection (bench / profiling) -------------------------------------------- */

/* cudaStream_t of the handle, as void*. */
int sfb_get_stream(sfb_sim* sim, void** stream);
/* Kernel launches issued by this handle so far (all kernels / hot-path kernels only). */
int sfb_get_launch_counts(sfb_sim* sim, int64_t* all_kernels, int64_t* step_kernels);
/* Per-kernel device time: while enabled every step records events around its three kernels
 * (k_sweep, k_rows, k_eval; a bitboard handle has no first kernel and reports k_tiles as the second).  sfb_get_kernel_ms returns the accumulated milliseconds and the
 * number of steps since enabling and resets them. */
int sfb_set_kernel_timing(sfb_sim* sim, int32_t enabled);
int sfb_get_kernel_ms(sfb_sim* sim, double* sweep_ms, double* rows_ms, double* eval_ms, int64_t* n_steps);
/* Row tasks (warp-rows that needed a cell-by-cell look) emitted by the last completed step. */
int sfb_get_row_tasks(sfb_sim* sim, int64_t* tasks, int64_t* capacity);
/* Units = (env, chunk of rows or single row, strip of columns) the last completed step listed,
 * and the number of units of the handle; mode: 0 = no unit skipping (every unit is swept, the two
 * counts are equal), 1 = flagged chunks are swept, 2 = flagged rows are the row tasks, 4 = listed tiles of 32 x 30 cells
 * (bitboard step: listed = tiles the last step looked at, total = tiles of the handle), 3 = list-driven step
 * (listed = entries of the watch list, total = cells of the handle). */
int sfb_get_unit_stats(sfb_sim* sim, int64_t* listed, int64_t* total, int32_t* mode);
/* Work-queue statistics of the last completed step: entries pushed, capacity, and
 * whether the step overflowed the queue and ran the dense fallback. */
int sfb_get_queue_stats(sfb_sim* sim, int64_t* entries, int64_t* capacity, int32_t* overflowed);
/* List handles (sfb_get_unit_stats mode 3): counters accumulated by k_front since the previous call, then
 * reset: stats[0] cells examined, [1] candidates evaluated (fire.py:163-234), [2] cells ignited, [3] sprites
 * pruned (fire.py:116-161), [4] cells that joined the watch list, [5] list entries read, [6] cells whose eight neighbours were read;
 * n <= 7 values.  Bitboard handles (mode 4) accumulate while kernel timing is on: [0] cells of the tiles
 * looked at, [1] candidates evaluated, [2] cells ignited, [3] sprites pruned, [5] tiles, [6] control-line cells
 * left to k_eval. */
int sfb_get_front_stats(sfb_sim* sim, int64_t* stats, int32_t n);
/* Test knob: enqueue a kernel that keeps the handle's stream busy for `microseconds` (at most 1 s)
 * before whatever is called next on it.  The GPU tests use it to widen the window of stream-ordering
 * races (a between-step call that is still running when the next call reads what it wrote). */
int sfb_debug_stall(sfb_sim* sim, int32_t microseconds);
/* Bytes of device memory held by the handle. */
int sfb_device_bytes(sfb_sim* sim, int64_t* bytes);

/* Rate of spread for n (direction, cell) pairs evaluated ON THE DEVICE with exactly the
 * code the step kernel uses -- the drop-in for compute_rate_of_spread
 * (simfire/world/rothermel.py:4-136).  dir[i] in 0..7 indexes the neighbour order of
 * fire.py:211-221; rec is float32 [n][8] in sfb_static_plane order; particle = h, S_T,
 * S_e, p_p, M_f; out is float64 [n] (ft/min, not yet scaled by update_rate). */
int sfb_rate_of_spread(int32_t device, const int8_t* dir, const float* rec, const float* particle,
                       int64_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif /* SIMFIRE_B200_H */
