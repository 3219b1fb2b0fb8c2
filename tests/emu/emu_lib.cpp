// TEST-ONLY translation unit: the product sources (C ABI + host glue + kernels), compiled with g++
// against the fiber emulator in cuda_emu.h.  Built into tests/emu/_build/ (git- and gpurun-ignored)
// by tests/emu/build_emu.py; never part of the package.
#define SFB_EMU 1
#define CUDA_EMU_IMPLEMENTATION
#include "cuda_emu.h"

#include "../../simfire_b200/csrc/sfb.cu"

extern "C" int sfb_emu_marker(void) { return 1; }
extern "C" long long sfb_emu_launches(void) { return emu::st().launches; }

// introspection for the emulator tests: the row tasks the last sweep emitted (whole-handle view
// or every group), as (env, y, strip) triples; returns the number of tasks
extern "C" long long sfb_emu_row_tasks(sfb_sim* s, int32_t* out, long long cap) {
    const int par = s->parity ^ 1;
    std::vector<EnvGroup*> views;
    if (s->last_mode == 2) for (auto& gr : s->groups) views.push_back(&gr);
    else views.push_back(&s->all);
    long long n = 0;
    for (EnvGroup* gr : views) {
        const long long cnt = (long long)gr->d.rows_count[par];
        const int e0 = (int)(gr->d.idx_base / s->d.plane);
        for (long long i = 0; i < cnt; ++i, ++n) {
            if (n >= cap) continue;
            const unsigned long long t = gr->d.rows[i];
            out[3 * n] = (int32_t)(t >> 28) + e0;
            out[3 * n + 1] = (int32_t)(t & 0xFFFFFu);
            out[3 * n + 2] = (int32_t)((t >> 20) & 0xFFu);
        }
    }
    return n;
}

// host-side microbenchmark hook (tests/emu/bench_patch.py): patch a synthetic change log into a
// mirror with the library's own apply_log on `threads` pool threads; returns milliseconds per call
extern "C" double sfb_emu_bench_apply_log(int H, int W, int E, const unsigned long long* log, long long n, int8_t* mirror,
                                          int threads, int reps, int single_step) {
    sfb_sim s{};
    s.d.H = H;
    s.d.W = W;
    s.d.E = E;
    s.d.pitch = (W + 15) / 16 * 16;
    s.d.plane = (int64_t)H * s.d.pitch;
    s.pool = new HostPool((unsigned)threads);
    s.patch_parallel_min = 16384;
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        const double t0 = now_ms();
        apply_log(&s, log, n, mirror, 0, (unsigned long long)E * s.d.plane, single_step != 0);
        best = std::min(best, now_ms() - t0);
    }
    delete s.pool;
    s.pool = nullptr;
    return best;
}
