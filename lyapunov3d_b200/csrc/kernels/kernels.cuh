// kernels.cuh -- the __global__ kernels (templated on exponent mode and sequence
// period) and the argument blocks they take by value.
#pragma once
#include <cuda_fp16.h>

#include "exponent.cuh"
#include "ray.cuh"

namespace lyap {

template <int MODE>
struct ArithOf { using type = ArithDev; };
template <>
struct ArithOf<kHost> { using type = ArithHost; };

// ----------------------------------------------------------------- volume bake
// Replaces kernel_calc_volume (reference kernel.cu:518-532).  The reference maps an
// 8x8x8 thread cube to voxels, so a warp stores four 32-byte fragments; here
// consecutive lanes own consecutive x, every warp store is one full 128-byte line
// (64 bytes for FP16), and a persistent grid strides over the slab.
struct BakeArgs {
    SeqPlan plan;
    float d;
    void *out;       // full volume
    int f16;         // 0: float, 1: __half
    uint32_t nx, ny, nz, z0, z1;
    unsigned long long *queue;   // next unclaimed work item of this launch (zeroed before launch)
};

template <int MODE>
__device__ __forceinline__ float voxel_coord(uint32_t i, uint32_t n)
{
    // a = 4.0f * (float)x / (float)N  (kernel.cu:527-529).  The reference's device build divides
    // with div.approx (kExact mirrors it: bit parity with the reference kernel on any grid);
    // its host build divides exactly (kHost, and kFast so that it tracks the CPU oracle on
    // grids that are not powers of two -- in chaotic voxels one ulp of coordinate moves l).
    if constexpr (MODE == kExact) return ArithDev::div(ArithDev::mul(__uint2float_rn(i), 4.0f), __uint2float_rn(n));
    else return __fdiv_rn(__fmul_rn(4.0f, __uint2float_rn(i)), __uint2float_rn(n));
}

__device__ __forceinline__ void store_voxel(const BakeArgs &a, uint64_t idx, float l)
{
    if (a.f16) reinterpret_cast<__half *>(a.out)[idx] = __float2half_rn(l);
    else reinterpret_cast<float *>(a.out)[idx] = l;
}

// Index arithmetic is 32-bit: lyap_bake() splits a slab into launches of < 2^31 voxels.
//
// Work is handed out dynamically, 32 consecutive items per warp per grab (one atomicAdd per
// ~10^4 cycles of work).  With a static grid-stride partition the warp scheduler's fixed
// priorities let some warps finish at half time and the pipes starve while the stragglers
// finish alone (ncu: 6.3 of 10 resident warps active on average, FMA pipe 83 % busy).
template <int MODE, int P>
__global__ void __launch_bounds__(256) bake_kernel(const __grid_constant__ BakeArgs a)
{
    typedef uint32_t IdxT;
    if constexpr (MODE == kHost) hostlog_init();
    const IdxT nx = a.nx;
    const IdxT plane = (IdxT)a.nx * (IdxT)a.ny;
    const IdxT total = plane * (IdxT)(a.z1 - a.z0);
    const uint64_t base = (uint64_t)a.z0 * a.nx * a.ny;
    const unsigned lane = threadIdx.x & 31;
    const IdxT items = (MODE == kFast) ? total / 2 + (total & 1) : total;   // fast mode: voxel pairs
    for (;;) {
        unsigned long long first = 0;
        if (lane == 0) first = atomicAdd(a.queue, 32ull);
        first = __shfl_sync(0xffffffffu, first, 0);
        if (first >= items) break;
        const IdxT k = (IdxT)first + lane;
        // lanes past the end of the slab evaluate a duplicate of the last item and skip the store
        const bool live = k < items;
        const IdxT kk = live ? k : items - 1;
        if constexpr (MODE == kFast) {
            // two x-adjacent voxels per lane through the packed evaluator: a warp covers 64
            // consecutive voxels and stores 256 contiguous bytes (FP32)
            const IdxT i0 = 2 * kk, i1 = (2 * kk + 1 < total) ? 2 * kk + 1 : 2 * kk;
            const IdxT q0 = i0 / plane, r0 = i0 - q0 * plane, y0 = r0 / nx, x0 = r0 - y0 * nx;
            IdxT q1 = q0, y1 = y0, x1 = x0;
            if (i1 != i0) {
                x1 = x0 + 1;
                if (x1 == nx) { x1 = 0; y1 = y0 + 1; if (y1 == (IdxT)a.ny) { y1 = 0; q1 = q0 + 1; } }
            }
            float l0, l1;
            exponent_fast2<P>(a.plan, voxel_coord<MODE>((uint32_t)x0, a.nx), voxel_coord<MODE>((uint32_t)y0, a.ny),
                              voxel_coord<MODE>(a.z0 + (uint32_t)q0, a.nz), voxel_coord<MODE>((uint32_t)x1, a.nx),
                              voxel_coord<MODE>((uint32_t)y1, a.ny), voxel_coord<MODE>(a.z0 + (uint32_t)q1, a.nz), a.d, l0, l1);
            if (live) {
                const uint64_t o0 = base + i0;
                // the pair goes out as one vector store only where its ADDRESS is vector-aligned: the
                // caller's pointer need only be element-aligned (a view into a larger allocation)
                const uintptr_t addr = reinterpret_cast<uintptr_t>(a.out) + o0 * (a.f16 ? 2u : 4u);
                if (!a.f16 && i1 != i0 && (addr & 7u) == 0) {
                    *reinterpret_cast<float2 *>(addr) = make_float2(l0, l1);
                } else if (a.f16 && i1 != i0 && (addr & 3u) == 0) {
                    *reinterpret_cast<__half2 *>(addr) = __floats2half2_rn(l0, l1);
                } else {
                    store_voxel(a, o0, l0);
                    if (i1 != i0) store_voxel(a, o0 + 1, l1);
                }
            }
        } else {
            const IdxT q = kk / plane, r = kk - q * plane, y = r / nx, x = r - y * nx;
            const float l = exponent<MODE, P>(a.plan, voxel_coord<MODE>((uint32_t)x, a.nx), voxel_coord<MODE>((uint32_t)y, a.ny),
                                              voxel_coord<MODE>(a.z0 + (uint32_t)q, a.nz), a.d);
            if (live) store_voxel(a, base + kk, l);
        }
    }
}

// ------------------------------------------------------- exponent at given points
struct PointsArgs {
    SeqPlan plan;
    float d;
    const float *xyz;
    float *out;
    uint64_t n;
};

template <int MODE, int P>
__global__ void __launch_bounds__(256) points_kernel(const __grid_constant__ PointsArgs a)
{
    if constexpr (MODE == kHost) hostlog_init();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride)
        a.out[i] = exponent<MODE, P>(a.plan, a.xyz[3 * i], a.xyz[3 * i + 1], a.xyz[3 * i + 2], a.d);
}

// ---------------------------------------------------------------------- render
// Replaces kernel_calc_render (reference kernel.cu:500-516): persistent warps pull
// pixels from a global queue; each lane carries one ray (ray.cuh) and the warp
// alternates between ONE uniform exponent evaluation for all 32 lanes and a short
// divergent state update.  Lanes whose ray finished are refilled with a single
// warp-aggregated atomic (ballot / popc / shfl).
struct RenderArgs {
    SeqPlan plan;
    lyap_cam cam;
    lyap_params prm;
    lyap_rgba *rgba;
    lyap_point *points;
    const lyap_light *lights;
    uint32_t n_lights;
    uint32_t width, height;
    uint32_t tile, tiles_x, n_tiles;   // tile edge, tiles per row, tiles in the image
    uint32_t rank, world;              // this launch renders tiles j with j % world == rank
    uint32_t compact;                  // write outputs densely in work order
    uint32_t nvcc_normal_quirk;        // test knob, see lyap_set_option("emulate_ref_nvcc_normals")
    unsigned long long n_items;        // work items of this rank: its tiles * tile^2
    unsigned long long *queue;         // next work item (zeroed before launch)
    unsigned long long *evals;         // optional: += exponent evaluations
    // hybrid mode (two launches on one stream).  The march kernel parks every ray whose
    // march ended inside the cube in the pixel's LyapPoint slot and appends its work item here; the
    // refine kernel (render_kernel with `worklist` set) takes its rays from the list.
    uint32_t *worklist;
    unsigned long long *work_count;
    float guard_scale;                 // test knob: factor on the per-sample guard band (1 = the derived bound, 0 = no band)
    uint32_t guard_batch;              // parked lanes per warp that trigger a parity-evaluator pass
    // volume-assisted march: a baked exponent volume classifies the cells of its grid;
    // march samples that fall into a cell whose whole neighbourhood is safely transparent are not
    // evaluated -- the ray just steps on (same positions, same arithmetic) and the cloud sums take the
    // cell's baked value.  Null `safe_bits` = plain hybrid march.
    const uint32_t *safe_bits;         // one bit per cell of the n^3 grid, x fastest
    const void *assist_vol;            // the baked volume the bits were built from (float or __half)
    uint32_t assist_n, assist_f16;
    float assist_scale;                // n / 4: sample coordinate -> cell index
    unsigned long long *skipped;       // optional: += skipped samples
    uint32_t sm_count;                 // SMs of the device (informational)
    uint32_t tail_compaction;          // 0 switches the tail protocol of render_kernel off
    // Optional permutation of the image's tiles (n_tiles entries): queue position -> tile.  The launch ends with
    // its longest ray, so the queue hands out the tiles whose rays CAN be long first and the provably short ones
    // last: tiles sorted by the length of the centre ray's chord through the cube, which bounds its march from
    // above (misc.cu: launch_tile_order).  Ranks interleave over queue positions, so every rank computes the
    // same permutation and takes positions rank, rank + world, ...  Null = image order.
    const uint32_t *tile_order;
};

constexpr int kRenderThreads = 128;
static_assert(kRenderThreads <= kSymTabThreads, "exponent_fast2 keeps one shared-memory column per thread");

// Work item k of a rank -> pixel.  Returns false for the padding of ragged edge tiles.
__device__ __forceinline__ bool item_to_pixel(const RenderArgs &a, unsigned long long k, uint32_t &x, uint32_t &y)
{
    const uint32_t tt = a.tile * a.tile;
    const uint32_t m = (uint32_t)(k / tt), p = (uint32_t)(k % tt);
    const uint32_t pos = m * a.world + a.rank;
    if (pos >= a.n_tiles) { x = y = 0; return false; }
    const uint32_t j = a.tile_order ? __ldg(a.tile_order + pos) : pos;
    x = (j % a.tiles_x) * a.tile + p % a.tile;
    y = (j / a.tiles_x) * a.tile + p / a.tile;
    return x < a.width && y < a.height;
}

template <class A>
__device__ __forceinline__ void finish_pixel(const RenderArgs &a, const RayState &st, bool hit)
{
    lyap_point pt;
    if (hit) {
        pt.P.x = st.Px; pt.P.y = st.Py; pt.P.z = st.Pz;
        pt.N.x = st.Nx; pt.N.y = st.Ny; pt.N.z = st.Nz;
        pt.a = st.a; pt.c = st.c; pt.l = st.l;
        a.points[st.out] = pt;
    } else {
        pt = a.points[st.out];   // a miss shades whatever the caller left there (kernel.cu:508-512)
    }
    reinterpret_cast<uint32_t *>(a.rgba)[st.out] = shade_pixel<A>(pt, a.cam, a.lights, a.n_lights);
}

// Tail compaction.  The unit of SFU / FMA work is a warp-instruction: a warp with one live ray costs its
// scheduler as much per evaluation as a full one.  Once the queue is empty the live rays thin out over
// all resident warps, and with four warps per scheduler every remaining ray advances at a quarter of
// the speed it would have alone -- the launch then ends 20-26 ms after its longest ray starts instead
// of ~7 ms (one 1 576-evaluation ray at full speed).  So from the moment the queue is drained the four
// warps of a block meet at a barrier once per evaluation, count their live rays, and whenever the rays
// fit into fewer warps they are repacked through shared memory into the first warps of an order that
// leaves the survivors of co-resident blocks on different schedulers (tail_rank); the emptied warps exit.  Rays are independent and their state is moved verbatim: results do not change.
struct TailShared {
    unsigned rank_tab[4];   // see tail_rank()
    int cnt[2][4];          // live rays per warp, double-buffered by iteration parity
    RayState pool[96];      // repacking buffer: a repack happens only when the rays fit into <= 3 warps
};

// How a warp that needs no ray learns that the queue is empty (it must then join its block's tail
// protocol): it looks at the queue head itself once per evaluation.  The load is issued before the
// evaluation and its value looked at after it, ~4 us later, so it costs no stall; the head only grows,
// so a stale value just means joining one evaluation later (the first barrier of the protocol waits).
__device__ __forceinline__ unsigned long long queue_peek(const unsigned long long *queue)
{
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(queue));
    return v;
}

__device__ __forceinline__ void block_barrier(int warps)
{
    asm volatile("bar.sync 1, %0;" ::"r"(warps * 32) : "memory");
}

// A warp's rank in its block's repacking order (rank r survives as long as the block needs more than r
// warps).  What matters is that the survivors of the blocks sharing an SM sit on DIFFERENT schedulers:
// repacked onto the same one they would gain nothing.  The scheduler of a warp is its hardware slot
// %warpid & 3 and a block's place on its SM is (%warpid >> 2) & 3 (measured on B200 with
// tools/probe_placement.cu: the block in place s gets the slots 4s + ((w + s) & 3) for its warps
// w = 0..3, so the hardware already staggers them; blockIdx says nothing reliable about the place).
// rank = scheduler - place makes the block in place s keep schedulers s, s+1, ... -- whatever the
// in-block numbering is.  `rank_tab` (shared, 4 entries) is used to check that the four ranks of the
// block are a cyclic shift of the warp index, which the pool layout relies on; otherwise rank = warp.
__device__ __forceinline__ unsigned tail_rank(unsigned *rank_tab, unsigned warp, unsigned lane)
{
    unsigned hw;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(hw));
    const unsigned rank = ((hw & 3u) - ((hw >> 2) & 3u)) & 3u;
    if (lane == 0) rank_tab[warp] = rank;
    __syncthreads();
    const unsigned shift = rank_tab[0];
    bool cyclic = true;
#pragma unroll
    for (unsigned w = 1; w < 4; ++w) cyclic = cyclic && (((rank_tab[w] - w) & 3u) == shift);
    return cyclic ? rank : warp;
}

// Four blocks of four warps per SM.  The exact evaluator wants the registers for its software pipeline (lg2
// results three steps away from their use).  The host evaluator was tried at 5 and 6 blocks (80-96 registers):
// no gain (1 244 vs 1 223 ms per 1080p frame) -- its step is bound by dispatch (18 instructions, of which
// the six FP64 ones hold the dispatch port for two cycles each), not by latency, once four warps share a
// scheduler.
template <int MODE, int P>
__global__ void __launch_bounds__(kRenderThreads, 4) render_kernel(const __grid_constant__ RenderArgs a)
{
    using A = typename ArithOf<MODE>::type;
    __shared__ TailShared ts;
    if constexpr (MODE == kHost) hostlog_init();

    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned rot = tail_rank(ts.rank_tab, warp, lane);   // this warp's rank in the repacking order (syncs the block)
    RayState st;
    st.phase = kNeedRay;
    st.sx = st.sy = st.sz = 3.0f;   // idle lanes evaluate a harmless dummy point (not 2.0: that orbit is superstable)
    bool drained = false, solo = false;
    int team = kRenderThreads / 32;   // warps of the block still taking part in the tail protocol
    unsigned it = 0;
    unsigned long long evals = 0, head = 0;   // head: the queue head as of the previous evaluation
    // hybrid mode's second launch: the rays are the ones the march kernel listed
    const unsigned long long n_items = a.worklist ? *a.work_count : a.n_items;

    for (;;) {
        if (head >= n_items) drained = true;
        // ---- refill idle lanes from the queue
        for (;;) {
            const bool need = st.phase == kNeedRay;
            const unsigned m = __ballot_sync(full, need);
            if (m == 0 || drained) break;
            const int leader = __ffs(m) - 1;
            const unsigned n_need = __popc(m);
            unsigned long long base = 0;
            if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)n_need);
            base = __shfl_sync(full, base, leader);
            if (need) {
                unsigned long long k = base + __popc(m & ((1u << lane) - 1u));
                uint32_t px, py;
                if (k < n_items) {
                    if (a.worklist) k = a.worklist[k];
                    if (item_to_pixel(a, k, px, py)) {
                        st.out = a.compact ? (uint32_t)k : px + py * a.width;
                        if (a.worklist) ray_resume<A>(st, px, py, a.cam, a.prm, a.points[st.out]);
                        else if (!ray_begin<A>(st, px, py, a.cam, a.prm)) finish_pixel<A>(a, st, false);
                    }
                }
            }
            if (base + n_need >= n_items) drained = true;
        }
        bool active = st.phase != kNeedRay;
        const unsigned live_mask = __ballot_sync(full, active);

        if (solo || !a.tail_compaction) {
            if (live_mask == 0 && drained) break;
        } else if (drained) {
            // ---- tail protocol: every warp of the team gets here once per evaluation
            const unsigned buf = it++ & 1u;
            if (lane == 0) ts.cnt[buf][warp] = __popc(live_mask);
            block_barrier(team);
            int total = 0, off = 0, live_warps = 0;
#pragma unroll
            for (unsigned r = 0; r < 4; ++r) {          // in repacking order; ranks >= team have left
                const int c = (int)r < team ? ts.cnt[buf][(r + warp - rot) & 3u] : 0;
                total += c;
                live_warps += c > 0;
                if (r < rot) off += c;
            }
            if (total == 0) break;
            const int want = (total + 31) / 32;
            if (want < live_warps || live_warps < team) {
                // repack the block's live rays into the first `want` warps of the order; the rest leave
                if (active) ts.pool[off + __popc(live_mask & ((1u << lane) - 1u))] = st;
                block_barrier(team);
                if ((int)rot >= want) break;
                const int idx = (int)rot * 32 + (int)lane;
                if (idx < total) {
                    st = ts.pool[idx];
                    active = true;
                } else {
                    st.phase = kNeedRay;
                    st.sx = st.sy = st.sz = 3.0f;
                    active = false;
                }
                team = want;
                // the pool is reused by a later repack only after the next count barrier, which every
                // remaining warp passes after it has read its rays here
                if (want == 1) solo = true;
            }
        }

        // ---- one exponent for every lane (idle lanes evaluate a dummy point)
        if (!drained) head = queue_peek(a.queue);
        const float l = exponent<MODE, P>(a.plan, st.sx, st.sy, st.sz, a.prm.d);

        // ---- advance each ray by that one sample
        if (active) {
            ++evals;
            const RayEvent ev = ray_advance<A>(st, l, a.prm);
            if (ev != kContinue) {
                if (ev == kHit) {
                    if (a.nvcc_normal_quirk) {
                        // What the reference's CUDA build computes under nvcc 12.9: ls[] shares its
                        // stack slot with lyap4d's abcd[], so ls[0..3] end up holding the last
                        // sample point and d (DESIGN.md "reference build defect").  Test knob only.
                        st.Nx = A::sub(st.Py, st.Px);
                        st.Ny = A::sub(a.prm.d, A::add(st.mag, st.Pz));
                    }
                    normalize3<A>(st.Nx, st.Ny, st.Nz);
                }
                finish_pixel<A>(a, st, ev == kHit);
                st.phase = kNeedRay;
            }
        }
    }

    if (a.evals) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(full, evals, o);
        if (lane == 0) atomicAdd(a.evals, evals);
    }
}

// The same tail protocol for the kernels that carry SLOTS rays per lane (a warp holds 32 * SLOTS rays).
template <class State, int SLOTS>
struct TailPool {
    unsigned rank_tab[4];
    int cnt[2][4];
    State pool[96 * SLOTS];
};

struct TailCtl {
    bool drained = false, solo = false;
    int team = kRenderThreads / 32;
    unsigned it = 0;
    unsigned long long head = 0;   // the queue head as of the previous evaluation (queue_peek)
};

// One turn of the protocol, called after the refill of every iteration.  live[j]: slot j holds a ray.
// Returns false when this warp has to leave the kernel's loop; may replace the warp's rays (and live[])
// by a repacked set.  `idle(state)` parks an empty slot.
template <class State, int SLOTS, class Idle>
__device__ __forceinline__ bool tail_turn(TailPool<State, SLOTS> &ts, TailCtl &c, State (&st)[SLOTS], bool (&live)[SLOTS],
                                          unsigned lane, unsigned warp, unsigned rot, bool enabled, Idle idle)
{
    const unsigned full = 0xffffffffu;
    unsigned mask[SLOTS];
    int mine = 0;
#pragma unroll
    for (int j = 0; j < SLOTS; ++j) {
        mask[j] = __ballot_sync(full, live[j]);
        mine += __popc(mask[j]);
    }
    if (c.solo || !enabled) return !(mine == 0 && c.drained);
    if (!c.drained) return true;
    const unsigned buf = c.it++ & 1u;
    if (lane == 0) ts.cnt[buf][warp] = mine;
    block_barrier(c.team);
    int total = 0, off = 0, live_warps = 0;
#pragma unroll
    for (unsigned r = 0; r < 4; ++r) {          // in repacking order; ranks >= team have left
        const int n = (int)r < c.team ? ts.cnt[buf][(r + warp - rot) & 3u] : 0;
        total += n;
        live_warps += n > 0;
        if (r < rot) off += n;
    }
    if (total == 0) return false;
    constexpr int kCap = 32 * SLOTS;
    const int want = (total + kCap - 1) / kCap;
    if (want < live_warps || live_warps < c.team) {
        int before = off;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            if (live[j]) ts.pool[before + __popc(mask[j] & ((1u << lane) - 1u))] = st[j];
            before += __popc(mask[j]);
        }
        block_barrier(c.team);
        if ((int)rot >= want) return false;
#pragma unroll
        for (int j = 0; j < SLOTS; ++j) {
            const int idx = (int)rot * kCap + j * 32 + (int)lane;
            if (idx < total) {
                st[j] = ts.pool[idx];
                live[j] = true;
            } else {
                idle(st[j]);
                live[j] = false;
            }
        }
        c.team = want;
        if (want == 1) c.solo = true;
    }
    return true;
}

// Fast mode: TWO rays per lane, evaluated together by the packed (f32x2) exponent.  The
// single-ray fast kernel is issue-bound (92 % of issue slots busy for 75 % FMA-pipe use,
// profiles/r01_render_fast_1080p.md); FFMA2/FMUL2 carry two rays' steps in one issue slot.
template <int P>
__global__ void __launch_bounds__(kRenderThreads, 3) render_fast2_kernel(const __grid_constant__ RenderArgs a)
{
    using A = ArithDev;
    __shared__ TailPool<RayState, 2> ts;
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned rot = tail_rank(ts.rank_tab, warp, lane);
    RayState st[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        st[j].phase = kNeedRay;
        st[j].sx = st[j].sy = st[j].sz = 3.0f;   // harmless dummy point for idle slots (2.0 would be the superstable orbit)
    }
    TailCtl tc;
    bool &drained = tc.drained;
    unsigned long long evals = 0;

    for (;;) {
        if (tc.head >= a.n_items) drained = true;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            for (;;) {
                const bool need = st[j].phase == kNeedRay;
                const unsigned m = __ballot_sync(full, need);
                if (m == 0 || drained) break;
                const int leader = __ffs(m) - 1;
                const unsigned n_need = __popc(m);
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)n_need);
                base = __shfl_sync(full, base, leader);
                if (need) {
                    const unsigned long long k = base + __popc(m & ((1u << lane) - 1u));
                    uint32_t px, py;
                    if (k < a.n_items && item_to_pixel(a, k, px, py)) {
                        st[j].out = a.compact ? (uint32_t)k : px + py * a.width;
                        if (!ray_begin<A>(st[j], px, py, a.cam, a.prm)) finish_pixel<A>(a, st[j], false);
                    }
                }
                if (base + n_need >= a.n_items) drained = true;
            }
        }
        bool live[2] = {st[0].phase != kNeedRay, st[1].phase != kNeedRay};
        if (!tail_turn(ts, tc, st, live, lane, warp, rot, a.tail_compaction != 0, [](RayState &s) {
                s.phase = kNeedRay;
                s.sx = s.sy = s.sz = 3.0f;
            }))
            break;

        float l[2];
        if (!drained) tc.head = queue_peek(a.queue);
        exponent_fast2<P>(a.plan, st[0].sx, st[0].sy, st[0].sz, st[1].sx, st[1].sy, st[1].sz, a.prm.d, l[0], l[1]);

#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (st[j].phase != kNeedRay) {
                ++evals;
                const RayEvent ev = ray_advance<A>(st[j], l[j], a.prm);
                if (ev != kContinue) {
                    if (ev == kHit) normalize3<A>(st[j].Nx, st[j].Ny, st[j].Nz);
                    finish_pixel<A>(a, st[j], ev == kHit);
                    st[j].phase = kNeedRay;
                    // park the slot on a harmless point: if it is never refilled it keeps being
                    // evaluated, and a stale zero-derivative point would send the whole warp
                    // through the safe re-evaluation on every remaining iteration
                    st[j].sx = st[j].sy = st[j].sz = 3.0f;
                }
            }
        }
    }

    if (a.evals) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(full, evals, o);
        if (lane == 0) atomicAdd(a.evals, evals);
    }
}

// ---------------------------------------------------------------- hybrid mode, first launch
// with prm.jitter == 0 a march sample's exponent is only ever COMPARED with the three
// thresholds (kernel.cu:326,370-384), so the march -- 97.7 % of all evaluations -- can run on the
// packed fast evaluator, two rays per lane, while refinement and normals (whose exponents end up
// in the LyapPoint record and in the pixel) stay on the parity evaluator MODE (kExact or kHost).
// A fast exponent that lies inside a guard band around a threshold -- sized per sample from a bound on
// the rounding error of the parity evaluator's float summation (fast_finish in exponent.cuh) -- could
// compare differently from the parity evaluator's: that slot is parked and the warp re-evaluates its parked samples with
// exponent<MODE> once `guard_batch` lanes hold one (or nothing else is left to do).  Hit point,
// normal, exponent and pixel are therefore those of the parity mode; only the cloud sums a and c
// (sums of the march exponents themselves) carry the fast evaluator's last-bit differences.
template <class A>
__device__ __forceinline__ void finish_miss(const RenderArgs &a, uint32_t out)
{
    const lyap_point pt = a.points[out];   // a miss shades whatever the caller left there (kernel.cu:508-512)
    reinterpret_cast<uint32_t *>(a.rgba)[out] = shade_pixel<A>(pt, a.cam, a.lights, a.n_lights);
}

template <int MODE, int P>
__global__ void __launch_bounds__(kRenderThreads, 3) march_fast2_kernel(const __grid_constant__ RenderArgs a)
{
    using A = typename ArithOf<MODE>::type;
    __shared__ TailPool<MarchState, 2> ts;
    if constexpr (MODE == kHost) hostlog_init();
    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31;
    const unsigned warp = threadIdx.x >> 5;
    const unsigned rot = tail_rank(ts.rank_tab, warp, lane);
    MarchState st[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        st[j].phase = kNeedRay;
        st[j].Px = st[j].Py = st[j].Pz = 3.0f;   // harmless dummy point for idle slots
    }
    TailCtl tc;
    bool &drained = tc.drained;
    unsigned long long evals = 0, skipped = 0;

    auto out_of = [&](uint32_t item) {
        uint32_t px, py;
        item_to_pixel(a, item, px, py);
        return a.compact ? item : px + py * a.width;
    };
    auto retire = [&](MarchState &s) {
        s.phase = kNeedRay;
        s.Px = s.Py = s.Pz = 3.0f;
    };
    // consume one march exponent of slot s (fast one outside the guard bands, or the parity one)
    auto consume = [&](MarchState &s, float l) {
        RayEvent ev = march_consume<A>(s, l, a.prm);
        if (ev == kContinue && a.safe_bits) {
            // volume assist: step over samples in safely transparent cells without evaluating them.
            // Only in the far-step state: a safe cell's exponents are above the near threshold, so an
            // evaluated sample could only ever switch near -> far there, never the other way.
            while (ev == kContinue && !s.near) {
                const float fx = s.Px * a.assist_scale, fy = s.Py * a.assist_scale, fz = s.Pz * a.assist_scale;
                const int ix = __float2int_rd(fx), iy = __float2int_rd(fy), iz = __float2int_rd(fz);
                const int n = (int)a.assist_n;
                if (!(ix >= 0 && ix < n && iy >= 0 && iy < n && iz >= 0 && iz < n)) break;    // NaN lands here too
                const uint32_t cell = (uint32_t)ix + a.assist_n * ((uint32_t)iy + a.assist_n * (uint32_t)iz);
                if (!((__ldg(a.safe_bits + (cell >> 5)) >> (cell & 31u)) & 1u)) break;
                const float lv = a.assist_f16 ? __half2float(__ldg(reinterpret_cast<const __half *>(a.assist_vol) + cell))
                                              : __ldg(reinterpret_cast<const float *>(a.assist_vol) + cell);
                march_account<A>(s, lv, a.prm);     // cloud sums from the baked value; cannot end the march or switch the step
                s.l = lv;
                ++skipped;
                ev = march_step<A>(s, a.prm);
            }
        }
        if (ev == kContinue) return;
        const uint32_t out = out_of(s.item);
        if (ev == kMiss) {
            finish_miss<A>(a, out);
        } else {
            lyap_point rec;
            rec.P.x = s.Px; rec.P.y = s.Py; rec.P.z = s.Pz;
            rec.N.x = s.t; rec.N.y = s.dt; rec.N.z = 0.0f;
            rec.a = s.a; rec.c = s.c; rec.l = s.l;
            a.points[out] = rec;
            a.worklist[atomicAdd(a.work_count, 1ull)] = s.item;
        }
        retire(s);
    };
    // g: the sample's own bound on the parity evaluator's summation error (fast_finish), plus that
    // evaluator's error per term -- lg2.approx is good to 2^-22 (1 + |log2 d|), glibc's logf to an ulp
    // of a term of magnitude <= 17
    constexpr float kTermEps = (MODE == kHost) ? 2e-6f : 1e-6f;
    auto in_guard = [&](float l, float g) {
        const float w = (g + kTermEps) * a.guard_scale;
        // NaN lands here too: the parity evaluator decides
        return !(fabsf(l - a.prm.opaqueThreshold) >= w && fabsf(l - a.prm.chaosThreshold) >= w && fabsf(l - a.prm.nearThreshold) >= w);
    };

    for (;;) {
        if (tc.head >= a.n_items) drained = true;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            for (;;) {
                const bool need = st[j].phase == kNeedRay;
                const unsigned m = __ballot_sync(full, need);
                if (m == 0 || drained) break;
                const int leader = __ffs(m) - 1;
                const unsigned n_need = __popc(m);
                unsigned long long base = 0;
                if ((int)lane == leader) base = atomicAdd(a.queue, (unsigned long long)n_need);
                base = __shfl_sync(full, base, leader);
                if (need) {
                    const unsigned long long k = base + __popc(m & ((1u << lane) - 1u));
                    uint32_t px, py;
                    if (k < a.n_items && item_to_pixel(a, k, px, py)) {
                        st[j].item = (uint32_t)k;
                        if (!ray_begin<A>(st[j], px, py, a.cam, a.prm)) {
                            finish_miss<A>(a, a.compact ? (uint32_t)k : px + py * a.width);
                            retire(st[j]);
                        }
                    }
                }
                if (base + n_need >= a.n_items) drained = true;
            }
        }
        {
            bool live[2] = {st[0].phase != kNeedRay, st[1].phase != kNeedRay};
            if (!tail_turn(ts, tc, st, live, lane, warp, rot, a.tail_compaction != 0, [](MarchState &s) {
                    s.phase = kNeedRay;
                    s.Px = s.Py = s.Pz = 3.0f;
                }))
                break;
        }
        const bool fast0 = st[0].phase == kMarchFirst || st[0].phase == kMarch;
        const bool fast1 = st[1].phase == kMarchFirst || st[1].phase == kMarch;
        const bool park0 = (st[0].phase & kGuardBit) != 0, park1 = (st[1].phase & kGuardBit) != 0;
        const bool any_fast = __ballot_sync(full, fast0 || fast1) != 0;
        const unsigned parked = __popc(__ballot_sync(full, park0 || park1));
        if (!any_fast && parked == 0) continue;   // (only before the queue is known to be empty; the tail protocol ends the loop)

        if (!drained) tc.head = queue_peek(a.queue);
        if (any_fast && parked < a.guard_batch) {
            // a parked slot sits the fast pass out on the dummy point: its own sample may well be a
            // zero-derivative one (NaN parks too), which would drag the warp through the fast
            // evaluator's out-of-line safe loop on every pass until the parity pass comes
            float l[2], g[2];
            exponent_fast2<P>(a.plan, park0 ? 3.0f : st[0].Px, park0 ? 3.0f : st[0].Py, park0 ? 3.0f : st[0].Pz,
                              park1 ? 3.0f : st[1].Px, park1 ? 3.0f : st[1].Py, park1 ? 3.0f : st[1].Pz, a.prm.d, l[0], l[1], &g[0], &g[1]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if (j == 0 ? fast0 : fast1) {
                    ++evals;
                    if (in_guard(l[j], g[j])) st[j].phase |= kGuardBit;
                    else consume(st[j], l[j]);
                }
            }
        } else {
            // one parked sample per lane through the parity evaluator (slot 0 first)
            const float x = park0 ? st[0].Px : st[1].Px, y = park0 ? st[0].Py : st[1].Py, z = park0 ? st[0].Pz : st[1].Pz;
            const float l = exponent<MODE, P>(a.plan, x, y, z, a.prm.d, 8);
            if (park0) { st[0].phase &= ~kGuardBit; consume(st[0], l); }
            else if (park1) { st[1].phase &= ~kGuardBit; consume(st[1], l); }
        }
    }

    if (a.evals) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) evals += __shfl_xor_sync(full, evals, o);
        if (lane == 0) atomicAdd(a.evals, evals);
    }
    if (a.skipped) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) skipped += __shfl_xor_sync(full, skipped, o);
        if (lane == 0) atomicAdd(a.skipped, skipped);
    }
}

// Shade-only pass over stored points (no marching): one thread per point.
struct ShadeArgs {
    lyap_cam cam;
    lyap_rgba *rgba;
    const lyap_point *points;
    const lyap_light *lights;
    uint32_t n_lights;
    unsigned long long count;
};

template <int MODE>
__global__ void __launch_bounds__(256) shade_kernel(const __grid_constant__ ShadeArgs a)
{
    using A = typename ArithOf<MODE>::type;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.count; i += stride)
        reinterpret_cast<uint32_t *>(a.rgba)[i] = shade_pixel<A>(a.points[i], a.cam, a.lights, a.n_lights);
}

// Classify the cells of a baked volume for the volume-assisted march: cell (i,j,k) spans the samples
// i..i+1 on each axis; it is SAFE when every sample of the neighbourhood i-dilate .. i+1+dilate exists,
// is finite and lies strictly between `lo` and `hi`.
struct AssistBuildArgs {
    uint32_t *bits;
    const void *vol;
    uint32_t f16, n, dilate;
    float lo, hi;
};
__global__ void __launch_bounds__(256) assist_build_kernel(const __grid_constant__ AssistBuildArgs a);

// Place one rank's compact (work-order) buffer into the full image.
struct ScatterArgs {
    uint8_t *image;
    const uint8_t *compact;
    uint32_t elem, width, height, tile, tiles_x, n_tiles, rank, world;
    unsigned long long n_items;
};

__global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ ScatterArgs a);

} // namespace lyap
