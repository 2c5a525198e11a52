// abi.cu -- the extern "C" GPU entry points of include/lyap/abi.h.
//
// Host-side work per call: validate, turn the reference's -1-terminated sequence into
// a SeqPlan (iteration schedule + which period instantiation to launch), size a
// persistent grid from the SM count and the kernel's occupancy, launch on the
// caller's stream.  No synchronisation, no allocation on the device entry points
// apart from a once-per-device scratch block of work-queue counters.
#include <cuda_runtime.h>

#include <time.h>

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "kernels/launch.hpp"
#include "lyap/abi.h"

using namespace lyap;

namespace {

// ------------------------------------------------------------------ options
std::atomic<long> g_force_generic{0};
std::atomic<long> g_render_warps_per_sm{0};   // 0 = default
std::atomic<long> g_bake_blocks_per_sm{0};    // 0 = occupancy maximum
std::atomic<long> g_fast_redo{1};             // debug knob: 0 disables the safe re-evaluation in fast mode
std::atomic<long> g_nvcc_normal_quirk{0};     // test knob: emulate the reference CUDA build's aliased normals
std::atomic<long> g_tile_order{1};            // frame queue in chord-descending tile order (0 = image order, for measurements)
std::atomic<long> g_tail_compaction{1};       // render_kernel's tail protocol (0 = off, for measurements)
std::atomic<long> g_guard_batch{0};           // hybrid mode: parked lanes per warp that trigger a parity pass (0 = default)
std::atomic<long> g_guard_scale{100};         // hybrid mode: guard band width in percent of the derived bound (test knob)
std::atomic<long> g_seq_table{1};             // generic periods: per-lane multiplier table in shared memory (0 = run-length loop)


// ---------------------------------------------------------------- sequence plan
const int kPeriods[] = {
#define X(p) p,
    LYAP_PERIODS(X)
#undef X
};

int pick_period(uint32_t len)
{
    if (g_force_generic.load()) return 0;
    int best = 0;
    for (int p : kPeriods)
        if (p > 0 && (uint32_t)p % len == 0 && (best == 0 || p < best)) best = p;
    return best;
}

// Builds the plan; returns the period instantiation to launch (0 = generic) or -1.
int build_plan(SeqPlan &sp, const int32_t *seq, uint32_t settle, uint32_t accum)
{
    if (!seq) return -1;
    uint32_t len = 0;
    while (seq[len] != -1) {
        if (seq[len] < 0 || seq[len] > 3) return -1;
        if (++len > LYAP_MAX_SEQUENCE) return -1;
    }
    if (len == 0) return -1;
    // the shortest period that generates the same infinite symbol stream
    uint32_t per = len;
    for (uint32_t p = 1; p < len; ++p) {
        if (len % p) continue;
        bool ok = true;
        for (uint32_t i = p; i < len && ok; ++i) ok = seq[i] == seq[i - p];
        if (ok) { per = p; break; }
    }
    const int P = pick_period(per);
    const uint32_t L = P > 0 ? (uint32_t)P : per;
    memset(&sp, 0, sizeof sp);
    sp.len = L;
    sp.fold_bias = (uint32_t)(127 + Accum<kFast>::kBias) << 23;
    sp.redo = (uint32_t)g_fast_redo.load();
    sp.settle = settle;
    sp.accum = accum;
    for (uint32_t i = 0; i < L; ++i) sp.sym[i] = (uint8_t)seq[i % per];
    // run-length form of the same period (runs split at 255; a run never wraps around the period end)
    sp.n_runs = 0;
    for (uint32_t i = 0; i < L;) {
        uint32_t j = i;
        while (j < L && sp.sym[j] == sp.sym[i] && j - i < 255) ++j;
        sp.runs[2 * sp.n_runs] = sp.sym[i];
        sp.runs[2 * sp.n_runs + 1] = (uint8_t)(j - i);
        ++sp.n_runs;
        i = j;
    }
    // Generic path: a per-lane multiplier table in shared memory costs one LDS per step, the run-length
    // loop ~25 instructions per run.  Measured on B200 (profiles/r02_longseq.log): runs of ~20 (A9A8B9B9)
    // 0.83 against 0.75 of the SFU roofline in favour of the run-length loop, runs of 10 (A9B9C9D9) 0.71
    // against 0.79 in favour of the table, an irregular 53-symbol sequence 0.25 against 0.76.  seq_table:
    // 0 = never, 1 = where runs average below 14 steps (default), 2 = always.  The launchers drop the table
    // again where its strips do not fit (launch.hpp).
    const long tab = g_seq_table.load();
    sp.table_stride = (P == 0 && (tab >= 2 || (tab == 1 && L < 14u * sp.n_runs))) ? (L | 1u) : 0u;
    sp.settle_head = settle % L;
    sp.settle_periods = settle / L;
    sp.accum_periods = accum / L;
    sp.accum_tail = accum % L;
    for (uint32_t k = 0; k < L && k < (uint32_t)kMaxPeriodRegs; ++k) sp.rot[k] = sp.sym[(sp.settle_head + k) % L];
    // symbol census over the accumulate steps (positions settle .. settle+accum-1)
    uint32_t per_period[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < L; ++i) per_period[sp.sym[i]]++;
    for (int s = 0; s < 4; ++s) sp.cnt[s] = per_period[s] * sp.accum_periods;
    for (uint32_t k = 0; k < sp.accum_tail; ++k) sp.cnt[sp.sym[(sp.settle_head + k) % L]]++;
    for (int s = 0; s < 4; ++s) sp.cnt_d[s] = (double)sp.cnt[s];
    sp.ln2_over_accum = 0.6931471805599453 / (double)accum;   // accum == 0: inf, and 0 * inf = NaN as before
    return P;
}

// The reference hands its kernels `cudaSeq`, a DEVICE copy of the -1-terminated array
// (lyap_interactive.cu:98-109, lyap_calculate.cu:45-56).  The entry points take either: a host
// (or managed) pointer is read in place; a device pointer is fetched with a small blocking
// copy, so that a caller can swap the <<<>>> line alone and keep passing cudaSeq.
struct SeqRef {
    const int32_t *ptr = nullptr;
    std::vector<int32_t> fetched;
    cudaError_t err = cudaSuccess;
    explicit SeqRef(const int32_t *seq)
    {
        ptr = seq;
        if (!seq) return;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, seq) != cudaSuccess) { cudaGetLastError(); return; }   // plain host memory on old drivers
        if (at.type != cudaMemoryTypeDevice) return;
        // length unknown: fetch a chunk; if that runs off the end of the allocation, go element by element
        const size_t chunk = 64;
        fetched.assign(LYAP_MAX_SEQUENCE + 2, -1);
        size_t have = 0;
        bool single = false;
        while (have <= LYAP_MAX_SEQUENCE) {
            const size_t n = single ? 1 : chunk;
            const size_t want = have + n <= fetched.size() ? n : fetched.size() - have;
            cudaError_t e = cudaMemcpy(fetched.data() + have, seq + have, want * sizeof(int32_t), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) {
                cudaGetLastError();
                if (single) { err = e; break; }
                single = true;
                continue;
            }
            bool done = false;
            for (size_t i = have; i < have + want; ++i) if (fetched[i] == -1) { done = true; break; }
            have += want;
            if (done) break;
        }
        fetched.back() = -1;
        ptr = fetched.data();
    }
};

// ------------------------------------------------------------ per-device scratch
struct DeviceScratch {
    int sm_count = 0;
    unsigned long long *counters = nullptr;   // ring of work-queue heads
    std::atomic<unsigned> next{0};
};
constexpr unsigned kCounterRing = 4096;
std::mutex g_mu;
DeviceScratch *g_scratch[64] = {};

cudaError_t scratch_for_current_device(DeviceScratch **out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lock(g_mu);
    if (!g_scratch[dev]) {
        DeviceScratch *s = new DeviceScratch;
        if ((e = cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) { delete s; return e; }
        if ((e = cudaMalloc(&s->counters, sizeof(unsigned long long) * kCounterRing)) != cudaSuccess) { delete s; return e; }
        if ((e = cudaMemset(s->counters, 0, sizeof(unsigned long long) * kCounterRing)) != cudaSuccess) { delete s; return e; }
        // hybrid mode takes its per-launch work list from the stream-ordered allocator: keep freed
        // blocks in the pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            cudaGetLastError();
        }
        g_scratch[dev] = s;
    }
    *out = g_scratch[dev];
    return cudaSuccess;
}

template <class F1, class F2, class F3>
auto by_mode(int mode, F1 exact, F2 fast, F3 host) -> decltype(exact())
{
    return mode == LYAP_MODE_FAST ? fast() : (mode == LYAP_MODE_HOST ? host() : exact());
}

bool valid_mode(int mode) { return mode >= LYAP_MODE_EXACT && mode <= LYAP_MODE_HYBRID_HOST; }
bool hybrid_mode(int mode) { return mode == LYAP_MODE_HYBRID || mode == LYAP_MODE_HYBRID_HOST; }
// The evaluator whose results a mode reproduces (what bake / points / shade run for the hybrids).
int parity_mode(int mode) { return mode == LYAP_MODE_HYBRID ? LYAP_MODE_EXACT : (mode == LYAP_MODE_HYBRID_HOST ? LYAP_MODE_HOST : mode); }

} // namespace

extern "C" {

const char *lyap_version(void) { return "lyapunov3d_b200 0.1 (sm_100a)"; }

const char *lyap_error_string(int code)
{
    switch (code) {
    case LYAP_OK: return "ok";
    case LYAP_ERR_BAD_SEQUENCE: return "bad sequence (empty, symbol outside 0..3, or too long)";
    case LYAP_ERR_BAD_ARGUMENT: return "bad argument";
    case LYAP_ERR_NO_DEVICE: return "no CUDA device";
    case LYAP_ERR_IO: return "i/o error";
    default: return cudaGetErrorString((cudaError_t)code);
    }
}

int lyap_set_option(const char *key, long value)
{
    if (!strcmp(key, "force_generic")) g_force_generic = value;
    else if (!strcmp(key, "seq_table")) g_seq_table = value;
    else if (!strcmp(key, "render_warps_per_sm")) g_render_warps_per_sm = value;
    else if (!strcmp(key, "bake_blocks_per_sm")) g_bake_blocks_per_sm = value;
    else if (!strcmp(key, "emulate_ref_nvcc_normals")) g_nvcc_normal_quirk = value;
    else if (!strcmp(key, "fast_redo")) g_fast_redo = value;
    else if (!strcmp(key, "tail_compaction")) g_tail_compaction = value;
    else if (!strcmp(key, "tile_order")) g_tile_order = value;
    else if (!strcmp(key, "hybrid_guard_batch")) g_guard_batch = value;
    else if (!strcmp(key, "hybrid_guard_percent")) g_guard_scale = value;
    else return LYAP_ERR_BAD_ARGUMENT;
    return LYAP_OK;
}

// LYAP_OPTIONS="key=value,key=value" in the environment applies lyap_set_option() when the library
// is loaded: the knobs of a program that links the library but does not call lyap_set_option itself
// (the patched reference programs of integration/).
namespace {
struct EnvOptions {
    EnvOptions()
    {
        const char *env = getenv("LYAP_OPTIONS");
        if (!env) return;
        std::string s(env);
        size_t pos = 0;
        while (pos < s.size()) {
            size_t end = s.find(',', pos);
            if (end == std::string::npos) end = s.size();
            const std::string kv = s.substr(pos, end - pos);
            const size_t eq = kv.find('=');
            if (eq != std::string::npos && lyap_set_option(kv.substr(0, eq).c_str(), atol(kv.c_str() + eq + 1)) != LYAP_OK)
                fprintf(stderr, "liblyap_b200: unknown option in LYAP_OPTIONS: %s\n", kv.c_str());
            pos = end + 1;
        }
    }
} g_env_options;
} // namespace

/* Which period instantiation a sequence would run on (0 = generic loop, -1 = invalid). */
int lyap_plan_period(const int32_t *seq, uint32_t settle, uint32_t accum)
{
    SeqPlan sp;
    return build_plan(sp, seq, settle, accum);
}

/* Test hook: the iteration schedule a sequence is turned into.  header[0..10] = period
 * instantiation, len, settle_head, settle_periods, accum_periods, accum_tail, cnt[0..3], n_runs;
 * sym/rot/runs receive the arrays (sym: len bytes, rot: min(len,32), runs: 2*n_runs). */
int lyap_plan_describe(const int32_t *seq, uint32_t settle, uint32_t accum, uint32_t *header, uint8_t *sym, uint8_t *rot, uint8_t *runs)
{
    SeqPlan sp;
    const int P = build_plan(sp, seq, settle, accum);
    if (P < 0) return LYAP_ERR_BAD_SEQUENCE;
    const uint32_t h[11] = {(uint32_t)P, sp.len, sp.settle_head, sp.settle_periods, sp.accum_periods, sp.accum_tail,
                            sp.cnt[0], sp.cnt[1], sp.cnt[2], sp.cnt[3], sp.n_runs};
    memcpy(header, h, sizeof h);
    memcpy(sym, sp.sym, sp.len);
    memcpy(rot, sp.rot, sp.len < (uint32_t)kMaxPeriodRegs ? sp.len : (uint32_t)kMaxPeriodRegs);
    memcpy(runs, sp.runs, 2 * sp.n_runs);
    return LYAP_OK;
}

uint64_t lyap_tile_count(uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world)
{
    if (!tile || tile > LYAP_MAX_TILE || !world || rank >= world) return 0;
    const uint64_t tiles = (uint64_t)((width + tile - 1) / tile) * ((height + tile - 1) / tile);
    const uint64_t mine = tiles / world + (rank < tiles % world ? 1 : 0);
    return mine * tile * tile;
}

namespace {
struct Assist {
    const void *vol;
    int dtype;
    uint32_t n;
    const uint32_t *bits;
    unsigned long long *skipped;
};

int render_impl(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world,
                int compact, int mode, unsigned long long *d_evals, void *stream, const Assist *assist)
{
    if (!d_rgba || !d_points || !cam || !prm || !valid_mode(mode) || !width || !height || !tile || tile > LYAP_MAX_TILE || !world || rank >= world)
        return LYAP_ERR_BAD_ARGUMENT;
    if (num_lights > LYAP_MAX_LIGHTS || (num_lights && !d_lights)) return LYAP_ERR_BAD_ARGUMENT;
    if ((uint64_t)width * height > 0xffffffffull) return LYAP_ERR_BAD_ARGUMENT;
    RenderArgs a;
    const SeqRef sq(seq);
    if (sq.err != cudaSuccess) return (int)sq.err;
    const int P = build_plan(a.plan, sq.ptr, prm->settle, prm->accum);
    if (P < 0) return LYAP_ERR_BAD_SEQUENCE;
    DeviceScratch *sc = nullptr;
    cudaError_t e = scratch_for_current_device(&sc);
    if (e != cudaSuccess) return (int)e;

    a.cam = *cam;
    a.prm = *prm;
    a.rgba = d_rgba;
    a.points = d_points;
    a.lights = d_lights;
    a.n_lights = num_lights;
    a.width = width;
    a.height = height;
    a.tile = tile;
    a.tiles_x = (width + tile - 1) / tile;
    a.n_tiles = a.tiles_x * ((height + tile - 1) / tile);
    a.rank = rank;
    a.world = world;
    a.compact = compact ? 1u : 0u;
    a.nvcc_normal_quirk = g_nvcc_normal_quirk.load() ? 1u : 0u;
    a.n_items = lyap_tile_count(width, height, tile, rank, world);
    a.evals = d_evals;
    a.sm_count = (uint32_t)sc->sm_count;
    a.tail_compaction = g_tail_compaction.load() ? 1u : 0u;
    if (a.n_items == 0) return LYAP_OK;

    cudaStream_t s = (cudaStream_t)stream;
    a.queue = sc->counters + (sc->next.fetch_add(1) % kCounterRing);
    if ((e = cudaMemsetAsync(a.queue, 0, sizeof(unsigned long long), s)) != cudaSuccess) return (int)e;
    // Queue order: tiles whose rays can be long first (kernels.cuh: RenderArgs::tile_order).  The compact
    // (work-order) output of the gather path keeps image order: lyap_scatter_tiles places by it.
    a.tile_order = nullptr;
    struct StreamScratch {
        void *p = nullptr;
        cudaStream_t s;
        ~StreamScratch() { if (p) cudaFreeAsync(p, s); }   // after everything queued on s below
    } order_mem;
    order_mem.s = s;
    if (!compact && a.n_tiles > 1 && a.n_tiles <= (1u << 30) && g_tile_order.load()) {   // (the sort counts in int)
        if ((e = cudaMallocAsync(&order_mem.p, tile_order_scratch_bytes(a.n_tiles), s)) != cudaSuccess) return (int)e;
        if ((e = launch_tile_order(a, order_mem.p, &a.tile_order, s)) != cudaSuccess) return (int)e;
    }
    a.worklist = nullptr;
    a.work_count = nullptr;
    a.safe_bits = nullptr;
    a.assist_vol = nullptr;
    a.assist_n = a.assist_f16 = 0;
    a.assist_scale = 0.0f;
    a.skipped = nullptr;

    // Hybrid modes: with jitter the march exponents feed the jitter PRNG (kernel.cu:334-347) and
    // every sample must be the parity evaluator's, so the call is the parity mode itself.
    if (hybrid_mode(mode) && prm->jitter == 0.0f) {
        if (a.n_items > 0xffffffffull) return LYAP_ERR_BAD_ARGUMENT;
        const bool host = mode == LYAP_MODE_HYBRID_HOST;
        a.guard_scale = (float)((double)g_guard_scale.load() / 100.0);
        const long gb = g_guard_batch.load();
        a.guard_batch = gb > 0 ? (uint32_t)(gb > 32 ? 32 : gb) : 1u;
        if (assist) {
            a.safe_bits = assist->bits;
            a.assist_vol = assist->vol;
            a.assist_n = assist->n;
            a.assist_f16 = assist->dtype == LYAP_F16;
            a.assist_scale = (float)assist->n * 0.25f;
            a.skipped = assist->skipped;
        }
        // [count, padded to 16 bytes][work list]: stream-ordered, lives until the second launch is done
        unsigned char *mem = nullptr;
        if ((e = cudaMallocAsync(&mem, 16 + a.n_items * sizeof(uint32_t), s)) != cudaSuccess) return (int)e;
        do {
            if ((e = cudaMemsetAsync(mem, 0, 16, s)) != cudaSuccess) break;
            a.work_count = reinterpret_cast<unsigned long long *>(mem);
            a.worklist = reinterpret_cast<uint32_t *>(mem + 16);
            int per_sm = host ? march_blocks_per_sm_host(P, a.plan) : march_blocks_per_sm_exact(P, a.plan);
            if (per_sm <= 0) { e = cudaErrorLaunchOutOfResources; break; }
            long want_warps = g_render_warps_per_sm.load();
            if (want_warps > 0) {
                if ((want_warps * 32 + kRenderThreads - 1) / kRenderThreads < per_sm)
                    per_sm = (int)((want_warps * 32 + kRenderThreads - 1) / kRenderThreads);
            } else {
                // small shards: fewer ray slots, more rays per slot (see the single-launch rule below)
                const unsigned long long slots = 2ull * sc->sm_count * per_sm * kRenderThreads;
                if (per_sm > 2 && a.n_items < 6ull * slots) per_sm -= 1;
            }
            unsigned long long grid = (unsigned long long)sc->sm_count * per_sm;
            const unsigned long long max_useful = (a.n_items + 2 * kRenderThreads - 1) / (2 * kRenderThreads);
            if (grid > max_useful) grid = max_useful;
            e = host ? launch_march_host(P, a, (unsigned)grid, s) : launch_march_exact(P, a, (unsigned)grid, s);
            if (e != cudaSuccess) break;
            // second launch: refinement + normals + shading of the listed rays on the parity evaluator
            RenderArgs b = a;
            b.queue = sc->counters + (sc->next.fetch_add(1) % kCounterRing);
            if ((e = cudaMemsetAsync(b.queue, 0, sizeof(unsigned long long), s)) != cudaSuccess) break;
            per_sm = host ? render_blocks_per_sm_host(P, a.plan) : render_blocks_per_sm_exact(P, a.plan);
            if (per_sm <= 0) { e = cudaErrorLaunchOutOfResources; break; }
            grid = (unsigned long long)sc->sm_count * per_sm;
            const unsigned long long max2 = (a.n_items + kRenderThreads - 1) / kRenderThreads;
            if (grid > max2) grid = max2;
            e = host ? launch_render_host(P, b, (unsigned)grid, s) : launch_render_exact(P, b, (unsigned)grid, s);
        } while (0);
        const cudaError_t ef = cudaFreeAsync(mem, s);
        return (int)(e != cudaSuccess ? e : ef);
    }
    mode = parity_mode(mode);

    int per_sm = by_mode(mode, [&] { return render_blocks_per_sm_exact(P, a.plan); }, [&] { return render_blocks_per_sm_fast(P, a.plan); },
                         [&] { return render_blocks_per_sm_host(P, a.plan); });
    if (per_sm <= 0) return (int)cudaErrorLaunchOutOfResources;
    // Persistent warps: as many as the kernel's registers allow (4 blocks x 4 warps per SM for the exact
    // evaluator, up to 6 x 4 for the host one).  A shard too small to give every lane a handful of rays
    // used to want fewer warps (round 1: hand-tuned thresholds), because its launch ended with all
    // resident warps dragging a few rays each; render_kernel now repacks those rays into fewer warps
    // (tail compaction), so full occupancy is right at every shard size.  The knob stays for measurements.
    const long want_warps = g_render_warps_per_sm.load();
    if (want_warps > 0) {
        const int want_blocks = (int)((want_warps * 32 + kRenderThreads - 1) / kRenderThreads);
        if (want_blocks < per_sm) per_sm = want_blocks;
    } else if (mode != LYAP_MODE_FAST) {
        // A lane works through its rays one after the other, so when the queue runs dry every lane still
        // holds a partly done ray: a residue of about one mean ray per lane that can no longer be balanced.
        // With fewer than ~9 rays per lane that residue is > 10 % of the launch, and trading one block per
        // SM (a quarter of the latency hiding, -1 % throughput on a full frame) for a third more rays per
        // lane pays (measured on 1/4 and 1/8 of a 1080p frame: 74.5 -> 72.6 ms, 43.4 -> 41.6 ms).
        const unsigned long long lanes = (unsigned long long)sc->sm_count * per_sm * kRenderThreads;
        if (per_sm > 3 && a.n_items < 9ull * lanes) per_sm -= 1;
    } else {
        // the two-ray kernel holds 2 x 128 rays per block: the same trade at three blocks per SM (measured on 1/8
        // and 1/4 of a 1080p frame, profiles/r02_shard_warps.log: 27.8 -> 26.5 ms, 45.5 -> 44.8 ms)
        const unsigned long long slots = 2ull * sc->sm_count * per_sm * kRenderThreads;
        if (per_sm > 2 && a.n_items < 6ull * slots) per_sm -= 1;
    }
    unsigned long long grid = (unsigned long long)sc->sm_count * per_sm;
    const unsigned long long max_useful = (a.n_items + kRenderThreads - 1) / kRenderThreads;
    if (grid > max_useful) grid = max_useful;

    e = by_mode(mode, [&] { return launch_render_exact(P, a, (unsigned)grid, s); },
                [&] { return launch_render_fast(P, a, (unsigned)grid, s); },
                [&] { return launch_render_host(P, a, (unsigned)grid, s); });
    return (int)e;
}

} // namespace

int lyap_render_tiles(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                      const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                      uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world,
                      int compact, int mode, unsigned long long *d_evals, void *stream)
{
    return render_impl(d_rgba, d_points, cam, prm, seq, d_lights, num_lights, width, height, tile, rank, world, compact, mode,
                       d_evals, stream, nullptr);
}

uint64_t lyap_assist_bits_bytes(uint32_t n) { return (((uint64_t)n * n * n + 31) / 32) * 4; }

int lyap_assist_build(uint32_t *d_safe_bits, const void *d_volume, int dtype, uint32_t n, const lyap_params *prm,
                      float margin, float upper, uint32_t dilate, void *stream)
{
    if (!d_safe_bits || !d_volume || !prm || n < 2 || n > 2048 || dilate > 8 || (dtype != LYAP_F32 && dtype != LYAP_F16) || !(margin >= 0.0f))
        return LYAP_ERR_BAD_ARGUMENT;
    DeviceScratch *sc = nullptr;
    cudaError_t e = scratch_for_current_device(&sc);
    if (e != cudaSuccess) return (int)e;
    AssistBuildArgs a;
    a.bits = d_safe_bits;
    a.vol = d_volume;
    a.f16 = dtype == LYAP_F16;
    a.n = n;
    a.dilate = dilate;
    // a skipped sample must neither end the march (l > opaque) nor switch the step (l > near)
    const float floor_thr = prm->opaqueThreshold > prm->nearThreshold ? prm->opaqueThreshold : prm->nearThreshold;
    a.lo = floor_thr + margin;
    a.hi = upper;
    return (int)launch_assist_build(a, (unsigned)sc->sm_count * 8u, (cudaStream_t)stream);
}

int lyap_render_assisted(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                         const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                         uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world, int compact, int mode,
                         const void *d_volume, int dtype, uint32_t n, const uint32_t *d_safe_bits,
                         unsigned long long *d_evals, unsigned long long *d_skipped, void *stream)
{
    if (!hybrid_mode(mode) || !d_volume || !d_safe_bits || n < 2 || n > 2048 || (dtype != LYAP_F32 && dtype != LYAP_F16))
        return LYAP_ERR_BAD_ARGUMENT;
    const Assist as{d_volume, dtype, n, d_safe_bits, d_skipped};
    return render_impl(d_rgba, d_points, cam, prm, seq, d_lights, num_lights, width, height, tile, rank, world, compact, mode,
                       d_evals, stream, &as);
}

int lyap_render(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                uint32_t width, uint32_t height, int mode, unsigned long long *d_evals, void *stream)
{
    return lyap_render_tiles(d_rgba, d_points, cam, prm, seq, d_lights, num_lights, width, height, 8, 0, 1, 0, mode, d_evals, stream);
}

int lyap_scatter_tiles(void *d_image, const void *d_compact, uint32_t elem_size, uint32_t width, uint32_t height,
                       uint32_t tile, uint32_t rank, uint32_t world, void *stream)
{
    if (!d_image || !d_compact || !elem_size || !tile || tile > LYAP_MAX_TILE || !world || rank >= world) return LYAP_ERR_BAD_ARGUMENT;
    ScatterArgs a;
    a.image = (uint8_t *)d_image;
    a.compact = (const uint8_t *)d_compact;
    a.elem = elem_size;
    a.width = width;
    a.height = height;
    a.tile = tile;
    a.tiles_x = (width + tile - 1) / tile;
    a.n_tiles = a.tiles_x * ((height + tile - 1) / tile);
    a.rank = rank;
    a.world = world;
    a.n_items = lyap_tile_count(width, height, tile, rank, world);
    if (!a.n_items) return LYAP_OK;
    const unsigned grid = (unsigned)((a.n_items + 255) / 256 < 4096 ? (a.n_items + 255) / 256 : 4096);
    return (int)launch_scatter(a, grid, (cudaStream_t)stream);
}

int lyap_shade_points(lyap_rgba *d_rgba, const lyap_point *d_points, const lyap_cam *cam,
                      const lyap_light *d_lights, uint32_t num_lights, uint64_t count, int mode, void *stream)
{
    if (!d_rgba || !d_points || !cam || !valid_mode(mode) || num_lights > LYAP_MAX_LIGHTS) return LYAP_ERR_BAD_ARGUMENT;
    if (!count) return LYAP_OK;
    mode = parity_mode(mode);
    ShadeArgs a;
    a.cam = *cam;
    a.rgba = d_rgba;
    a.points = d_points;
    a.lights = d_lights;
    a.n_lights = num_lights;
    a.count = count;
    const unsigned grid = (unsigned)((count + 255) / 256 < 8192 ? (count + 255) / 256 : 8192);
    return (int)launch_shade(mode, a, grid, (cudaStream_t)stream);
}

int lyap_bake(void *d_exps, int dtype, const lyap_params *prm, const int32_t *seq,
              uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z0, uint32_t z1, int mode, void *stream)
{
    if (!d_exps || !prm || !valid_mode(mode) || !nx || !ny || !nz || z1 > nz || z0 > z1) return LYAP_ERR_BAD_ARGUMENT;
    if (dtype != LYAP_F32 && dtype != LYAP_F16) return LYAP_ERR_BAD_ARGUMENT;
    mode = parity_mode(mode);
    BakeArgs a;
    const SeqRef sq(seq);
    if (sq.err != cudaSuccess) return (int)sq.err;
    const int P = build_plan(a.plan, sq.ptr, prm->settle, prm->accum);
    if (P < 0) return LYAP_ERR_BAD_SEQUENCE;
    if (z0 == z1) return LYAP_OK;
    DeviceScratch *sc = nullptr;
    cudaError_t e = scratch_for_current_device(&sc);
    if (e != cudaSuccess) return (int)e;
    a.d = prm->d;
    a.out = d_exps;
    a.f16 = dtype == LYAP_F16;
    a.nx = nx; a.ny = ny; a.nz = nz; a.z0 = z0; a.z1 = z1;

    int per_sm = by_mode(mode, [&] { return bake_blocks_per_sm_exact(P, a.plan); }, [&] { return bake_blocks_per_sm_fast(P, a.plan); },
                         [&] { return bake_blocks_per_sm_host(P, a.plan); });
    if (per_sm <= 0) return (int)cudaErrorLaunchOutOfResources;
    const long cap = g_bake_blocks_per_sm.load();
    if (cap > 0 && cap < per_sm) per_sm = (int)cap;
    // the kernel indexes a launch's voxels with 32 bits: cut tall slabs into z-chunks
    const unsigned long long plane = (unsigned long long)nx * ny;
    if (plane >= (1ull << 31)) return LYAP_ERR_BAD_ARGUMENT;
    const uint32_t max_planes = (uint32_t)(((1ull << 31) - 1) / plane);
    cudaStream_t s = (cudaStream_t)stream;
    for (uint32_t za = z0; za < z1; za += max_planes) {
        const uint32_t zb = (z1 - za > max_planes) ? za + max_planes : z1;
        a.z0 = za;
        a.z1 = zb;
        a.queue = sc->counters + (sc->next.fetch_add(1) % kCounterRing);
        if ((e = cudaMemsetAsync(a.queue, 0, sizeof(unsigned long long), s)) != cudaSuccess) break;
        const unsigned long long total = plane * (zb - za);
        const unsigned long long per_thread = (mode == LYAP_MODE_FAST) ? 2 : 1;
        unsigned long long grid = (unsigned long long)sc->sm_count * per_sm;
        const unsigned long long need = (total / per_thread + 255) / 256 + 1;
        if (grid > need) grid = need;
        e = by_mode(mode, [&] { return launch_bake_exact(P, a, (unsigned)grid, s); }, [&] { return launch_bake_fast(P, a, (unsigned)grid, s); },
                    [&] { return launch_bake_host(P, a, (unsigned)grid, s); });
        if (e != cudaSuccess) break;
    }
    return (int)e;
}

int lyap_exponent_points(float *d_out, const float *d_xyz, uint64_t n, const lyap_params *prm,
                         const int32_t *seq, int mode, void *stream)
{
    if (!d_out || !d_xyz || !prm || !valid_mode(mode)) return LYAP_ERR_BAD_ARGUMENT;
    mode = parity_mode(mode);
    PointsArgs a;
    const SeqRef sq(seq);
    if (sq.err != cudaSuccess) return (int)sq.err;
    const int P = build_plan(a.plan, sq.ptr, prm->settle, prm->accum);
    if (P < 0) return LYAP_ERR_BAD_SEQUENCE;
    if (!n) return LYAP_OK;
    DeviceScratch *sc = nullptr;
    cudaError_t e = scratch_for_current_device(&sc);
    if (e != cudaSuccess) return (int)e;
    a.d = prm->d;
    a.xyz = d_xyz;
    a.out = d_out;
    a.n = n;
    unsigned long long grid = (unsigned long long)sc->sm_count * 4;
    if (grid > (n + 255) / 256) grid = (n + 255) / 256;
    cudaStream_t s = (cudaStream_t)stream;
    e = by_mode(mode, [&] { return launch_points_exact(P, a, (unsigned)grid, s); }, [&] { return launch_points_fast(P, a, (unsigned)grid, s); },
                [&] { return launch_points_host(P, a, (unsigned)grid, s); });
    return (int)e;
}

int lyap_ray_probe(float *d_out, const uint32_t *d_pixels, uint64_t n, const lyap_cam *cam, const lyap_params *prm, int mode, void *stream)
{
    if (!d_out || !d_pixels || !cam || !prm || !valid_mode(mode)) return LYAP_ERR_BAD_ARGUMENT;
    if (!n) return LYAP_OK;
    return (int)launch_ray_probe(parity_mode(mode) == LYAP_MODE_HOST ? kHost : kExact, d_out, d_pixels, n, *cam, *prm, (cudaStream_t)stream);
}

int lyap_normalize_vectors(float *d_xyz, uint64_t n, int mode, void *stream)
{
    if (!d_xyz || !valid_mode(mode)) return LYAP_ERR_BAD_ARGUMENT;
    if (!n) return LYAP_OK;
    return (int)launch_normalize(parity_mode(mode) == LYAP_MODE_HOST ? kHost : kExact, d_xyz, n, (cudaStream_t)stream);
}

// --------------------------------------------------------------- host-buffer path
// Per-device workspace of the host-buffer calls: device buffers and a stream that live across
// calls (cudaMalloc/cudaFree of ~80 MB per frame and their implicit synchronisations cost more
// than the copies).  One call at a time per device; lyap_host_workspace_release() frees it.
namespace {
struct HostWorkspace {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    void *buf[4] = {nullptr, nullptr, nullptr, nullptr};   // rgba, points, lights, evals
    size_t cap[4] = {0, 0, 0, 0};
};
HostWorkspace g_ws[64];

cudaError_t ws_reserve(HostWorkspace &w, int slot, size_t bytes)
{
    if (w.cap[slot] >= bytes) return cudaSuccess;
    if (w.buf[slot]) cudaFree(w.buf[slot]);
    w.buf[slot] = nullptr;
    w.cap[slot] = 0;
    const cudaError_t e = cudaMalloc(&w.buf[slot], bytes);
    if (e == cudaSuccess) w.cap[slot] = bytes;
    return e;
}

double now_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
} // namespace

void lyap_host_workspace_release(void)
{
    for (int d = 0; d < 64; ++d) {
        HostWorkspace &w = g_ws[d];
        std::lock_guard<std::mutex> lock(w.mu);
        bool any = w.stream != nullptr;
        for (int k = 0; k < 4; ++k) any = any || w.buf[k];
        if (!any) continue;
        cudaSetDevice(d);
        for (int k = 0; k < 4; ++k) {
            if (w.buf[k]) cudaFree(w.buf[k]);
            w.buf[k] = nullptr;
            w.cap[k] = 0;
        }
        if (w.stream) cudaStreamDestroy(w.stream);
        w.stream = nullptr;
    }
}

int lyap_render_host(lyap_rgba *h_rgba, lyap_point *h_points, const lyap_cam *cam, const lyap_params *prm,
                     const int32_t *seq, const lyap_light *h_lights, uint32_t num_lights,
                     uint32_t width, uint32_t height, int mode, int device, unsigned long long *evals_out)
{
    if (!h_rgba || !cam || !prm || !seq || (num_lights && !h_lights) || device < 0 || device >= 64) return LYAP_ERR_BAD_ARGUMENT;
    // reject before anything is reserved or copied: the light buffer below holds LYAP_MAX_LIGHTS entries
    if (num_lights > LYAP_MAX_LIGHTS || !valid_mode(mode) || !width || !height || (uint64_t)width * height > 0xffffffffull)
        return LYAP_ERR_BAD_ARGUMENT;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    const bool trace = getenv("LYAP_TRACE") != nullptr;   // stage timings on stderr (adds synchronisations)
    const size_t n = (size_t)width * height;
    HostWorkspace &w = g_ws[device];
    std::lock_guard<std::mutex> lock(w.mu);
    int rc = LYAP_OK;
    double t0 = now_ms(), t1 = t0, t2 = t0, t3 = t0;
    do {
        if (!w.stream && (e = cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = ws_reserve(w, 0, n * sizeof(lyap_rgba))) != cudaSuccess) break;
        if ((e = ws_reserve(w, 1, n * sizeof(lyap_point))) != cudaSuccess) break;
        if ((e = ws_reserve(w, 2, sizeof(lyap_light) * LYAP_MAX_LIGHTS)) != cudaSuccess) break;
        if ((e = ws_reserve(w, 3, sizeof(unsigned long long))) != cudaSuccess) break;
        cudaStream_t s = w.stream;
        lyap_rgba *d_rgba = (lyap_rgba *)w.buf[0];
        lyap_point *d_points = (lyap_point *)w.buf[1];
        lyap_light *d_lights = (lyap_light *)w.buf[2];
        unsigned long long *d_evals = (unsigned long long *)w.buf[3];
        // the reference never clears its point buffer; zero gives miss pixels a defined colour
        if ((e = cudaMemsetAsync(d_points, 0, n * sizeof(lyap_point), s)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(d_rgba, 0, n * sizeof(lyap_rgba), s)) != cudaSuccess) break;
        if ((e = cudaMemsetAsync(d_evals, 0, sizeof(unsigned long long), s)) != cudaSuccess) break;
        if (num_lights && (e = cudaMemcpyAsync(d_lights, h_lights, sizeof(lyap_light) * num_lights, cudaMemcpyHostToDevice, s)) != cudaSuccess) break;
        if (trace) { cudaStreamSynchronize(s); t1 = now_ms(); }
        rc = lyap_render(d_rgba, d_points, cam, prm, seq, d_lights, num_lights, width, height, mode, d_evals, s);
        if (rc != LYAP_OK) break;
        if (trace) { cudaStreamSynchronize(s); t2 = now_ms(); }
        if ((e = cudaMemcpyAsync(h_rgba, d_rgba, n * sizeof(lyap_rgba), cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        if (h_points && (e = cudaMemcpyAsync(h_points, d_points, n * sizeof(lyap_point), cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        unsigned long long ev = 0;
        if ((e = cudaMemcpyAsync(&ev, d_evals, sizeof ev, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) break;
        if (evals_out) *evals_out = ev;
        t3 = now_ms();
    } while (0);
    if (trace) fprintf(stderr, "[lyap trace] render_host %ux%u: setup+H2D %.3f ms, kernel %.3f ms, D2H %.3f ms\n", width, height, t1 - t0, t2 - t1, t3 - t2);
    if (rc != LYAP_OK) return rc;
    return (int)e;
}

int lyap_bake_host(void *h_exps, int dtype, const lyap_params *prm, const int32_t *seq,
                   uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z0, uint32_t z1, int mode, int device)
{
    if (!h_exps || !prm || !seq || !nx || !ny || !nz || z1 > nz || z0 > z1 || device < 0 || device >= 64) return LYAP_ERR_BAD_ARGUMENT;
    if (!valid_mode(mode) || (dtype != LYAP_F32 && dtype != LYAP_F16)) return LYAP_ERR_BAD_ARGUMENT;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int)e;
    const size_t esz = dtype == LYAP_F16 ? 2 : 4;
    const size_t plane = (size_t)nx * ny;
    const size_t bytes = plane * (z1 - z0) * esz;
    if (!bytes) return LYAP_OK;
    // the slab is contiguous in the volume: allocate just the slab and bias the base pointer
    char *d_slab = nullptr;
    cudaStream_t s = nullptr;
    int rc = LYAP_OK;
    do {
        if ((e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_slab, bytes)) != cudaSuccess) break;
        rc = lyap_bake(d_slab - plane * z0 * esz, dtype, prm, seq, nx, ny, nz, z0, z1, mode, s);
        if (rc != LYAP_OK) break;
        if ((e = cudaMemcpyAsync((char *)h_exps + plane * z0 * esz, d_slab, bytes, cudaMemcpyDeviceToHost, s)) != cudaSuccess) break;
        e = cudaStreamSynchronize(s);
    } while (0);
    cudaFree(d_slab);
    if (s) cudaStreamDestroy(s);
    if (rc != LYAP_OK) return rc;
    return (int)e;
}

// ------------------------------------------------------------------ peer memory
// One process per GPU: rank 0 allocates the result buffer, exports a CUDA IPC handle, the
// other ranks map it and hand the mapped pointer to lyap_bake / lyap_render_tiles, whose
// kernels then store their shard straight into rank 0's HBM over NVLink/NVSwitch.  No
// gather step exists: the "collective" is the kernels' own stores.
int lyap_peer_alloc(void **dptr, uint64_t bytes)
{
    if (!dptr || !bytes) return LYAP_ERR_BAD_ARGUMENT;
    cudaError_t e = cudaMalloc(dptr, bytes);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaMemset(*dptr, 0, bytes);
}

int lyap_peer_free(void *dptr) { return (int)cudaFree(dptr); }

int lyap_peer_export(void *dptr, unsigned char *handle64)
{
    if (!dptr || !handle64) return LYAP_ERR_BAD_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, dptr);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle64, &h, 64);
    return LYAP_OK;
}

int lyap_peer_open(const unsigned char *handle64, void **dptr)
{
    if (!dptr || !handle64) return LYAP_ERR_BAD_ARGUMENT;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int lyap_peer_close(void *dptr) { return (int)cudaIpcCloseMemHandle(dptr); }

int lyap_probe_ffma2(double *packed_ffma_lane_ops_per_s)
{
    double v = 0;
    const cudaError_t e = probe_ffma2(&v);
    if (e != cudaSuccess) return (int)e;
    if (packed_ffma_lane_ops_per_s) *packed_ffma_lane_ops_per_s = v;
    return LYAP_OK;
}

int lyap_probe_peaks(double *ffma_lane_ops_per_s, double *mufu_lane_ops_per_s, double *sm_clock_hz_est, int *sm_count)
{
    double f = 0, m = 0, c = 0;
    int n = 0;
    const cudaError_t e = probe_peaks(&f, &m, &c, &n);
    if (e != cudaSuccess) return (int)e;
    if (ffma_lane_ops_per_s) *ffma_lane_ops_per_s = f;
    if (mufu_lane_ops_per_s) *mufu_lane_ops_per_s = m;
    if (sm_clock_hz_est) *sm_clock_hz_est = c;
    if (sm_count) *sm_count = n;
    return LYAP_OK;
}

} // extern "C"
