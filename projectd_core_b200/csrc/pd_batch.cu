/*
 * pd_batch.cu -- CUDA kernels (sm_100a) + the C ABI of include/pd_batch.h.
 *
 * Kernel map (one thread = one car unless noted):
 *   k_tick          Simulator::step for every env: Car::step + stepComponents + dWorldStep + postStep
 *                   (device functions in pd_tick.h / pd_car.h / pd_solver.h / pd_track.h)
 *   k_teleport      Car::teleportToSpline (reset path)
 *   k_set_controls / k_set_actions   setCarControls / the env's action mapping
 *   k_observe       the 24-float observation of pyprojectd/projectd_env.py:237-275
 *   env_epilogue    (inside the tick kernels) observation + reward / termination logic of ProjectDEnv.step
 *                   (projectd_env.py:178-212) + episode statistics
 *   k_raycast       batch rays against the track BVH
 * State is structure-of-arrays in HBM (include/pd_state.h); car parameters and the track are read-only
 * device buffers shared by all envs (served from L2/L1 after first touch).
 * No CPU fallback exists: every entry point that computes launches a kernel or fails.
 */
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "pd_quad.h"
#include "pd_lbvh.h"
#include "host/pd_host.h"
#include "../../include/pd_batch.h"

using namespace pd;

#define PD_BLOCK 64
#ifndef PD_SERIAL_SMEM_SCRATCH
#define PD_SERIAL_SMEM_SCRATCH 0   /* thread-per-car kernel: D | dg part of the solver scratch in shared memory (1) or everything in local memory (0); measured equal at 65536 envs (0.884 ms), 0 is 8 % faster at 16384 */
#endif
#define PD_QBLOCK 64   /* threads per block of the quad kernel = stride of the lane-interleaved solver scratch */

/* ------------------------------------------------------------------ kernels ------------------------------------------------------------------ */
/* Optional env-step work fused into the tick kernels (null pointers: skipped).
 *   act      [n][2]  actions -> controls BEFORE the tick (projectd_env.py:158-170)
 *   obs      [n][24] observation AFTER the tick (:237-275)
 *   reward / done / envReturn / envLen / stats: ProjectDEnv.step tail (:178-212) + episode statistics */
struct EnvIO {
    const float* act; float* obs;
    float* reward; int32_t* done; float* envReturn; int32_t* envLen; double* stats;
    double timeAfter;
    /* PD_AUTORESET_NEXT_STEP: an env that finished at step t is reset INSIDE the tick kernel of step t+1
     * (teleport + the reset's zero-action tick, projectd_env.py:216-227), its action of that step is ignored */
    int32_t* pending;            /* [n] 1 = finished at the previous step (null: no in-kernel reset) */
    uint32_t* episodeCtr; uint64_t seed, idOffset; int teleportMode;
    const int32_t* collIn;       /* [n] k_collide's answer for this tick's start pose (0 / 1; -1 = test inside the tick); null: test inside the tick */
    const float* contacts;       /* [n][PD_CONTACT_WORDS] live contact joints per env (collision response on), or null */
    PdEnvConfig cfg;             /* ProjectDEnv's knobs (gas range, penalties, termination switches, clutch / gear overrides) */
    long long* clk;              /* profiling aid (PD_DEBUG_CLOCKS=1): SM cycles each warp spent in the tick, [blocks * 2]; null otherwise */
};

/* counter-based uniform in [0,1): splitmix64 of (seed, global env id, episode counter) */
__host__ __device__ inline float pd_uniform(uint64_t seed, uint64_t id, uint64_t ctr) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (id * 0x100000001B3ull + ctr + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

/* the reset of one env inside a tick kernel (one thread): teleportCarByMode + the zero action of ProjectDEnv.reset */
template <class SVX> __device__ __noinline__ void env_reset_in_kernel(const PdCarParams& P, const TrackDev& T, const SVX& sv, int e, const EnvIO& io, double time) {
    float u = 0.0f;
    if (io.teleportMode == PD_TELEPORT_NEAREST) u = sv.f(PD_OFF_CAR + PD_CAR_o_trackLocation);
    else if (io.teleportMode == PD_TELEPORT_RANDOM) { u = pd_uniform(io.seed, io.idOffset + (uint64_t)e, io.episodeCtr[e]); io.episodeCtr[e]++; }
    car_teleport_to_point(P, T, sv, point_id_at_distance(T, u), time);
    sv.i(PD_OFF_CAR + PD_CAR_o_nanFlag, 0);
    env_apply_action(sv, 0.0f, 0.0f, io.cfg);
}

template <class SVX> __device__ __forceinline__ void env_epilogue(const SVX& sv, int e, bool on, const EnvIO& io, unsigned warpMask, bool resetNow = false) {
    if (on && io.obs) {
        float o[PD_OBS_DIM];
        car_observe(sv, o);
        float4* dst = reinterpret_cast<float4*>(io.obs + (size_t)e * PD_OBS_DIM);     /* 96 B per env, 16-byte aligned */
#pragma unroll
        for (int k = 0; k < PD_OBS_DIM / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    }
    if (!io.reward) return;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (on && resetNow) {      /* the reset's own step: ProjectDEnv.reset returns the observation only and zeroes the episode counters */
        io.reward[e] = 0.0f; io.done[e] = 0; io.envReturn[e] = 0; io.envLen[e] = 0; io.pending[e] = 0;
    } else if (on) {
        float r; int d;
        env_reward_done(sv, io.timeAfter, io.cfg, r, d);
        const float ret = io.envReturn[e] + r; const int len = io.envLen[e] + 1;
        if (ret < io.cfg.terminate_low_reward) d |= PD_DONE_LOWREWARD;
        io.reward[e] = r; io.done[e] = d;
        if (d) {
            s[0] = 1; s[1] = ret; s[2] = len;
            s[3] = (d & PD_DONE_COLLISION) ? 1 : 0; s[4] = (d & PD_DONE_OFFTRACK) ? 1 : 0; s[5] = (d & PD_DONE_STUCK) ? 1 : 0;
            s[6] = (d & PD_DONE_LOWREWARD) ? 1 : 0; s[7] = (d & PD_DONE_NAN) ? 1 : 0;
            io.envReturn[e] = 0; io.envLen[e] = 0;
            if (io.pending) io.pending[e] = 1;
        } else { io.envReturn[e] = ret; io.envLen[e] = len; }
    }
    /* warp-level reduction, one atomic per warp and statistic */
    const bool any = __any_sync(warpMask, s[0] != 0.0);
    if (!any) return;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double v = s[k];
        for (int off = 16; off > 0; off >>= 1) { const double o2 = __shfl_down_sync(warpMask, v, off); if ((threadIdx.x & 31) + off < 32 && ((warpMask >> ((threadIdx.x & 31) + off)) & 1u)) v += o2; }
        if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(&io.stats[k], v);
    }
}

/* the tick, one thread per car, tiled structure-of-arrays state (batches above the quad threshold) */
#ifndef PD_SERIAL_MINBLOCKS
#define PD_SERIAL_MINBLOCKS 8
#endif
/* MB = resident blocks per SM the register budget is set for: 8 (128 registers) lets a 65536-env batch sit in ONE wave (1024 blocks of 64 threads on
 * 148 x 8 slots); a batch of at most 148 x 4 x 64 = 37888 envs fits one wave at 4 blocks per SM, where ptxas may take 255 registers and keeps more of the
 * tick out of local memory: measured +7.7 % at 24576 envs, +10.5 % at 32768 (and -45 % at 65536, where it means two waves). */
template <int TOPO, int MB = PD_SERIAL_MINBLOCKS>       /* TOPO: suspension topology (PD_TOPO_*), one compile-time instance per (front, rear) pair of the bundled cars */
#if defined(PD_SERIAL_MAXNREG)
__global__ void __maxnreg__(PD_SERIAL_MAXNREG) k_tick(      /* experiment: an explicit register cap instead of the launch bound.  144 registers (7 blocks of 64 threads per SM on paper) measured 62.4 vs 83.1 M car-ticks/s at 65536 envs: the allocation granularity leaves room for 6 blocks only, i.e. two waves */
#else
__global__ void __launch_bounds__(PD_BLOCK, MB) k_tick(
#endif
                                                   const __grid_constant__ PdCarParams P, const __grid_constant__ TrackDev T, uint32_t* state, int n, float dt, double time,
                                                   const int32_t* __restrict__ mask, const __grid_constant__ EnvIO io) {
    const long long clk0 = io.clk ? clock64() : 0;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const bool on = e < n && (!mask || mask[e]);
    SVTile sv = sv_tiled(state, (size_t)(e < n ? e : 0));
    /* solver scratch: the rows part (JA | JB | Y, 209 words) in local memory, the D | dg part (77 words, re-read by
       every phase of the factorisation) in shared memory, lane-interleaved: 19.7 KB per block keeps 8 blocks per SM */
#if PD_SOLVER2
    float* pd_rows = nullptr;      /* register-resident solver (pd_solver2.h): no scratch */
#elif PD_SERIAL_SMEM_SCRATCH
    __shared__ float pd_scrD[PD_GSCR_D_WORDS * PD_BLOCK];
    float pd_rows[PD_GSCR_ROWS_WORDS];
#else
    float pd_rows[PD_GSCR_WORDS];
#endif
    const bool resetNow = on && io.pending && io.pending[e];
    const int collPre = (on && io.collIn) ? io.collIn[e] : -1;
    if (on) {
        if (resetNow) env_reset_in_kernel(P, T, sv, e, io, time);
        else if (io.act) env_apply_action(sv, io.act[e * 2 + 0], io.act[e * 2 + 1], io.cfg);
#if PD_SOLVER2
        car_tick<1, 1, TOPO>(P, T, sv, dt, time, pd_rows, pd_rows, collPre, io.contacts ? io.contacts + (size_t)e * PD_CONTACT_WORDS : nullptr);
#elif PD_SERIAL_SMEM_SCRATCH
        car_tick<1, PD_BLOCK>(P, T, sv, dt, time, pd_rows, pd_scrD + threadIdx.x, collPre);
#else
        car_tick<1, 1>(P, T, sv, dt, time, pd_rows, pd_rows + PD_GSCR_ROWS_WORDS, collPre);
#endif
    }
    if (io.clk && (threadIdx.x & 31) == 0) io.clk[blockIdx.x * (PD_BLOCK / 32) + (threadIdx.x >> 5)] = clock64() - clk0;
    env_epilogue(sv, e, on, io, 0xffffffffu, resetNow);
}

/* exchange policy of pd_quad.h on the GPU: shuffles inside the quad, with the quad's own member mask */
struct QuadShfl {
    int lane; unsigned mask; int base;
    /* helper lanes: when a warp holds at most 4 cars, a second quad per car (lane + 4*CPW) runs the same tick on the same data and
       takes half of the work of the sections that split cleanly (half = 0 main / 1 helper, nhalf = 1 or 2); peer() swaps a value
       with the twin lane, mask8 covers both quads */
    int half, nhalf, off; unsigned mask8;
    __device__ __forceinline__ float peer(float v) const { return nhalf == 2 ? __shfl_xor_sync(mask8, v, off) : v; }
    __device__ __forceinline__ int peer(int v) const { return nhalf == 2 ? __shfl_xor_sync(mask8, v, off) : v; }
#if defined(PD_PHASE_CLOCKS)
    long long* ph;
#endif
    __device__ __forceinline__ float get(float v, int src) const { return __shfl_sync(mask, v, base + src); }
    __device__ __forceinline__ int get(int v, int src) const { return __shfl_sync(mask, v, base + src); }
    __device__ __forceinline__ V3 get(V3 v, int src) const { return v3(get(v.x, src), get(v.y, src), get(v.z, src)); }
    __device__ __forceinline__ float sum(float v) const { v += __shfl_xor_sync(mask, v, 1); v += __shfl_xor_sync(mask, v, 2); return v; }
    __device__ __forceinline__ V3 sum(V3 v) const { return v3(sum(v.x), sum(v.y), sum(v.z)); }
    __device__ __forceinline__ bool all(bool p) const { return (__ballot_sync(mask, p) & mask) == mask; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask8); }
};

/* ---- bulk asynchronous copies HBM <-> shared memory (the TMA engine's 1-D path: SASS UBLKCP) ---- */
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_addr(src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef PD_QUAD_LOCAL_SCRATCH
#define PD_QUAD_LOCAL_SCRATCH 0   /* solver scratch of the quad kernel: 0 = shared memory, 1 = local memory, 2 = JA | JB local, Y | D | dg shared */
#endif
/* shared memory of one block of the quad kernel with CPW cars per warp: 2*CPW records | mbarrier | scratch of 8*CPW lanes */
#define PD_QUAD_SMEM_BYTES_(CPW) (2 * (CPW) * PD_STATE_STRIDE * 4 + 16 + ((PD_SOLVER2 || PD_QUAD_LOCAL_SCRATCH == 1) ? 0 : PD_QUAD_LOCAL_SCRATCH == 2 ? 8 * (CPW) * PD_GSCR_D_WORDS * 4 : 8 * (CPW) * PD_GSCR_WORDS * 4))

/* the tick, four lanes per car, array-of-records state staged through shared memory:
 * block = 64 threads = 2 warps, CPW cars per warp (8: every lane busy, throughput; 4 or 2: half / quarter-filled warps
 * = more warps per car batch, for batches too small to fill the 592 warp schedulers of a B200 otherwise).
 * ONE bulk copy brings the block's records in, the quads work on the shared-memory copy, ONE bulk copy writes them back. */
template <int CPW, int TOPO = 0>
/* registers: left to ptxas (168 with the plain bound).  Measured on B200: capping at 128 (5 blocks per SM) is slower everywhere
 * (4096 envs 23.7 vs 26.4 M car-ticks/s, 65536 envs 33 vs 52 M); an explicit minimum of 1 block makes ptxas take 255 registers. */
#ifdef PD_QUAD_MINBLOCKS
__global__ void __launch_bounds__(PD_QBLOCK + 32, PD_QUAD_MINBLOCKS) k_tick_quad(
#else
__global__ void __launch_bounds__(PD_QBLOCK + 32) k_tick_quad(
#endif
                                                         const __grid_constant__ PdCarParams P, const __grid_constant__ TrackDev T, uint32_t* state, int n, float dt, double time,
                                                         const int32_t* __restrict__ mask, const __grid_constant__ EnvIO io) {
    constexpr int QCARS = 2 * CPW;          /* cars per block */
    constexpr int QLANES = 8 * CPW;         /* working threads per block = stride of the lane-interleaved solver scratch */
    extern __shared__ __align__(128) uint32_t pd_smem[];
    uint32_t* recs = pd_smem;                                                         /* [QCARS][PD_STATE_STRIDE] */
    uint64_t* bar = reinterpret_cast<uint64_t*>(pd_smem + QCARS * PD_STATE_STRIDE);
    float* scratch = reinterpret_cast<float*>(pd_smem + QCARS * PD_STATE_STRIDE + 4);  /* [PD_GSCR_WORDS][QLANES], lane-interleaved */
    const long long clk0 = io.clk ? clock64() : 0;
    const int tid = threadIdx.x;
    const int wl = tid & 31, warp = tid >> 5;
    constexpr bool HELP = (8 * CPW <= 32);                       /* room for a helper quad per car in the same warp */
    const bool worker = wl < (HELP ? 8 * CPW : 4 * CPW);         /* lanes beyond the warp's cars (and their helpers) only help with the block-level steps */
    const int cid = warp * (4 * CPW) + (worker ? (wl % (4 * CPW)) : 0);   /* compact index of a working lane (a helper shares its main lane's) */
    const int car0 = blockIdx.x * QCARS;
    const int ncars = min(QCARS, n - car0);
    const int car = cid >> 2, e = car0 + car;
    const bool on = worker && car < ncars && (!mask || mask[e]);
    if (mask && !__syncthreads_or(on)) return;                 /* reset pass: nothing to do for this block */
    const uint32_t bytes = (uint32_t)ncars * PD_STATE_STRIDE * 4;
    uint32_t* gsrc = state + (size_t)car0 * PD_STATE_STRIDE;
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (tid == 0) { mbar_expect_tx(bar, bytes); bulk_g2s(recs, gsrc, bytes, bar); }
    mbar_wait(bar, 0);
    uint32_t* rec = recs + car * PD_STATE_STRIDE;
    SVFlat sv = sv_flat(rec);
    const bool resetNow = on && io.pending && io.pending[e];
    QuadShfl ex; ex.lane = wl & 3; ex.base = wl & ~3; ex.mask = 0xFu << ex.base;
    ex.off = 4 * CPW; ex.nhalf = HELP ? 2 : 1; ex.half = (HELP && wl >= 4 * CPW) ? 1 : 0;
    ex.mask8 = HELP ? (ex.mask | (ex.half ? (ex.mask >> ex.off) : (ex.mask << ex.off))) : ex.mask;
#if defined(PD_PHASE_CLOCKS)
    ex.ph = (io.clk && wl == 0 && warp < 2) ? io.clk + 4096 + (size_t)(blockIdx.x * 2 + warp) * 32 : nullptr;   /* profiling build: 32 stamps per warp after the per-warp totals */
#endif
    if (on) {
        if (ex.lane == 0) rec[PD_OFF_TYRE(0) + PD_TYRE_o_tyrePad] = 0u;      /* the collision warp's mailbox (a pad word of the record) starts empty */
        if (resetNow) { if (ex.lane == 0 && ex.half == 0) env_reset_in_kernel(P, T, sv, e, io, time); ex.sync(); }
        else if (io.act) { env_apply_action(sv, io.act[e * 2 + 0], io.act[e * 2 + 1], io.cfg); ex.sync(); }   /* four identical writes */
    }
    /* A third warp, when launched (blockDim 96: full ticks without k_collide's answer), is the block's COLLISION WARP: it
     * takes the start poses of the block's cars (after a reset's teleport, before anything moves) into registers and tests the
     * cars on an odd physics frame one after the other with all 32 lanes (car_collide_warp) while the two tick warps run the
     * tick; the answer is posted in a pad word of the car's record, where the quad picks it up just before the scoring. */
    const bool haveCollWarp = blockDim.x > PD_QBLOCK && !io.collIn;
    const int collSlot = PD_OFF_TYRE(0) + PD_TYRE_o_tyrePad;
    if (haveCollWarp) {
        __syncthreads();
        float fr[12]; int need = 0;
        if (warp == 2) {
            if (wl < ncars && (!mask || mask[car0 + wl])) {
                SVFlat svc = sv_flat(recs + wl * PD_STATE_STRIDE);
                Body Cc; load_body(svc, PD_BODY_CHASSIS, Cc);
                fr[0] = Cc.fr.p.x; fr[1] = Cc.fr.p.y; fr[2] = Cc.fr.p.z; fr[3] = Cc.fr.ax.x; fr[4] = Cc.fr.ax.y; fr[5] = Cc.fr.ax.z;
                fr[6] = Cc.fr.ay.x; fr[7] = Cc.fr.ay.y; fr[8] = Cc.fr.ay.z; fr[9] = Cc.fr.az.x; fr[10] = Cc.fr.az.y; fr[11] = Cc.fr.az.z;
                need = svc.i(PD_OFF_CAR + PD_CAR_o_physFrame) & 1;
            } else { PD_UNROLL for (int q = 0; q < 12; ++q) fr[q] = 0.0f; }
        }
        __syncthreads();      /* poses and frame parities are in the collision warp's registers: the tick may start changing the records */
        if (warp == 2) {
#ifndef PD_COLL_WARP_SEQ
#define PD_COLL_WARP_SEQ 1      /* measured on B200 (30-step bench runs): sequential 23.6 / side by side 23.7 M car-ticks/s at 4096 envs, 8.65 / 8.53 M at 1024, 14.6 / 14.3 M at 2048, 36.6 / 37.0 M at 8192: no difference worth the code; the odd-frame cost is not the collision warp's own duration */
#endif
#if !PD_COLL_WARP_SEQ
            /* The block's cars side by side: 32 / QCARS lanes per car run the floor-box test (the cells of the footprint dealt to them, as in
               k_collide2) -- 93 % of the cars need nothing else (tools/collide_tail.py) -- then the whole warp takes the cars whose footprint holds
               WALL triangles at hull height through the hull test, one after the other.  (Round 1 tested the cars one after the other with all
               32 lanes: 8 x 15 k cycles, and a block with two or three kerb-straddling cars kept its quads waiting: odd frames 198-221 us
               against 167 us for even ones under ncu.  ONE lane per car was slower still: 22.4 vs 25.8 M car-ticks/s.) */
            constexpr int LPC = 32 / QCARS;
            const unsigned FULLW = 0xffffffffu;
            const int kc = wl / LPC, sub = wl % LPC;
            Body Cc;
            Cc.fr.p = v3(__shfl_sync(FULLW, fr[0], kc), __shfl_sync(FULLW, fr[1], kc), __shfl_sync(FULLW, fr[2], kc));
            Cc.fr.ax = v3(__shfl_sync(FULLW, fr[3], kc), __shfl_sync(FULLW, fr[4], kc), __shfl_sync(FULLW, fr[5], kc));
            Cc.fr.ay = v3(__shfl_sync(FULLW, fr[6], kc), __shfl_sync(FULLW, fr[7], kc), __shfl_sync(FULLW, fr[8], kc));
            Cc.fr.az = v3(__shfl_sync(FULLW, fr[9], kc), __shfl_sync(FULLW, fr[10], kc), __shfl_sync(FULLW, fr[11], kc));
            const bool needK = __shfl_sync(FULLW, need, kc) != 0;
            bool hitK = false, walls = false;
            if (needK) hitK = car_collide(P, T, Cc, 0, 1, sub, LPC, 1, &walls);
            PD_UNROLL
            for (int off = 1; off < LPC; off <<= 1) { hitK = __shfl_xor_sync(FULLW, (int)hitK, off) || hitK; walls = __shfl_xor_sync(FULLW, (int)walls, off) || walls; }
            if (hitK) walls = false;
            if (needK && !walls && sub == 0) { *reinterpret_cast<volatile uint32_t*>(recs + kc * PD_STATE_STRIDE + collSlot) = hitK ? 2u : 1u; __threadfence_block(); }      /* answered: this car's quad need not wait for the others */
            unsigned m = __ballot_sync(FULLW, walls && sub == 0);
            while (m) {
                const int k = (__ffs(m) - 1) / LPC; m &= m - 1;
                Body Cw;
                Cw.fr.p = v3(__shfl_sync(FULLW, fr[0], k), __shfl_sync(FULLW, fr[1], k), __shfl_sync(FULLW, fr[2], k));
                Cw.fr.ax = v3(__shfl_sync(FULLW, fr[3], k), __shfl_sync(FULLW, fr[4], k), __shfl_sync(FULLW, fr[5], k));
                Cw.fr.ay = v3(__shfl_sync(FULLW, fr[6], k), __shfl_sync(FULLW, fr[7], k), __shfl_sync(FULLW, fr[8], k));
                Cw.fr.az = v3(__shfl_sync(FULLW, fr[9], k), __shfl_sync(FULLW, fr[10], k), __shfl_sync(FULLW, fr[11], k));
                const bool hit = car_collide_warp<false, false>(P, T, Cw, wl, nullptr);
                if (wl == 0) { *reinterpret_cast<volatile uint32_t*>(recs + k * PD_STATE_STRIDE + collSlot) = hit ? 2u : 1u; __threadfence_block(); }
            }
#else
            for (int k = 0; k < ncars; ++k) {
                if (!__shfl_sync(0xffffffffu, need, k)) continue;
                Body Cc;
                Cc.fr.p = v3(__shfl_sync(0xffffffffu, fr[0], k), __shfl_sync(0xffffffffu, fr[1], k), __shfl_sync(0xffffffffu, fr[2], k));
                Cc.fr.ax = v3(__shfl_sync(0xffffffffu, fr[3], k), __shfl_sync(0xffffffffu, fr[4], k), __shfl_sync(0xffffffffu, fr[5], k));
                Cc.fr.ay = v3(__shfl_sync(0xffffffffu, fr[6], k), __shfl_sync(0xffffffffu, fr[7], k), __shfl_sync(0xffffffffu, fr[8], k));
                Cc.fr.az = v3(__shfl_sync(0xffffffffu, fr[9], k), __shfl_sync(0xffffffffu, fr[10], k), __shfl_sync(0xffffffffu, fr[11], k));
                const bool hit = car_collide_warp<false>(P, T, Cc, wl, nullptr);
                if (wl == 0) { *reinterpret_cast<volatile uint32_t*>(recs + k * PD_STATE_STRIDE + collSlot) = hit ? 2u : 1u; __threadfence_block(); }
            }
#endif
        }
    }
    if (on) {
        const int collPre = io.collIn ? io.collIn[e] : -1;
        volatile uint32_t* collWait = haveCollWarp ? reinterpret_cast<volatile uint32_t*>(rec + collSlot) : nullptr;
#if PD_QUAD_LOCAL_SCRATCH == 1
        float lscr[PD_GSCR_WORDS];
        car_tick_quad<1, 1>(P, T, sv, dt, time, ex, lscr, lscr + PD_GSCR_ROWS_WORDS, collPre, collWait);
#elif PD_QUAD_LOCAL_SCRATCH == 2
        float lrows[PD_GSCR_ROWS_WORDS];                     /* JA | JB in local memory, Y | D | dg in shared memory */
        car_tick_quad<1, QLANES>(P, T, sv, dt, time, ex, lrows, scratch + cid, collPre, collWait);
#else
        car_tick_quad<QLANES, QLANES, TOPO>(P, T, sv, dt, time, ex, scratch + cid, scratch + PD_GSCR_ROWS_WORDS * QLANES + cid, collPre, collWait, io.contacts ? io.contacts + (size_t)e * PD_CONTACT_WORDS : nullptr);
#endif
    }
    if (io.clk && wl == 0 && warp < 2) io.clk[blockIdx.x * 2 + warp] = clock64() - clk0;
    fence_async_smem();                                        /* generic-proxy writes -> visible to the bulk copy engine */
    __syncthreads();
    if (tid == 32) { bulk_s2g(gsrc, recs, bytes); }
    if (tid < 32) {                                            /* warp 0: one thread per car of the block */
        const int c2 = tid, e2 = car0 + c2;
        const bool on2 = c2 < ncars && (!mask || mask[e2]);
        SVFlat sv2 = sv_flat(recs + (c2 < QCARS ? c2 : 0) * PD_STATE_STRIDE);
        const bool reset2 = on2 && io.pending && io.pending[e2];     /* read before the epilogue clears it; the quads read it before the block barrier above */
        env_epilogue(sv2, e2, on2, io, 0xffffffffu, reset2);
    }
    if (tid == 32) bulk_wait_all();                            /* the copy engine has read (and written) everything before the block retires */
}

/* Collision detection for the coming tick, ONE WARP PER CAR: the cells of the car's footprint are dealt to
 * 4 groups of 8 lanes, the entries of a cell's lists to the 8 lanes of a group (pd_collide.h), the 32 answers are OR-ed.
 * Runs ahead of the tick kernel on the same stream, on the tick's start pose (the pose collisionStep sees,
 * PhysicsEngineODE.cpp:216-224); envs on an even physics frame answer 0 at once; for envs about to be reset inside the tick
 * kernel the pose is the one the teleport will produce. */
#define PD_COLLIDE_BLOCK 128
#ifndef PD_COLLIDE_MINBLOCKS
#define PD_COLLIDE_MINBLOCKS 4      /* measured on B200 at 65536 envs: 8 (64 registers, 32 warps per SM) is SLOWER than 4 (128 registers): 71.6 vs 79.1 M car-ticks/s */
#endif
__global__ void __launch_bounds__(PD_COLLIDE_BLOCK, PD_COLLIDE_MINBLOCKS) k_collide(const __grid_constant__ PdCarParams P, const __grid_constant__ TrackDev T, const uint32_t* __restrict__ state, int layout, int n,
                                                              const int32_t* __restrict__ pending, int32_t* __restrict__ collOut, long long* __restrict__ dbg,
                                                              int teleportMode, uint64_t seed, uint64_t idOffset, const uint32_t* __restrict__ episodeCtr) {
    const long long clk0 = dbg ? clock64() : 0;
    const int e = (blockIdx.x * PD_COLLIDE_BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n) return;
    SVR sv = sv_env(layout, const_cast<uint32_t*>(state), (size_t)e);
    if (!(sv.i(PD_OFF_CAR + PD_CAR_o_physFrame) & 1)) { if (lane == 0) collOut[e] = 0; return; }
    Body C; load_body(sv, PD_BODY_CHASSIS, C);
    if (pending && pending[e]) {
        /* this env is reset inside the coming tick kernel, BEFORE its collision step: test the pose the teleport will give it
           (same spline point: the episode counter is only read here, the tick kernel advances it) */
        float u = 0.0f;
        if (teleportMode == PD_TELEPORT_NEAREST) u = sv.f(PD_OFF_CAR + PD_CAR_o_trackLocation);
        else if (teleportMode == PD_TELEPORT_RANDOM) u = pd_uniform(seed, idOffset + (uint64_t)e, episodeCtr[e]);
        Quat q; teleport_chassis_pose(P, T, point_id_at_distance(T, u), C.fr.ax, C.fr.ay, C.fr.az, q, C.fr.p);
    }
    int stats[6] = {0, 0, 0, 0, 0, 0};
    __shared__ __align__(16) float hullS[(PD_COLLIDE_BLOCK / 32) * PD_HULLS_WORDS];
    const bool any = car_collide_warp<true>(P, T, C, lane, hullS + (threadIdx.x >> 5) * PD_HULLS_WORDS, dbg ? stats : nullptr);
    if (lane == 0) collOut[e] = any ? 1 : 0;
    if (dbg && lane == 0) { dbg[4096 + (size_t)n * 12 + (size_t)e * 4] = clock64() - clk0; dbg[4096 + (size_t)n * 12 + (size_t)e * 4 + 1] = ((long long)stats[0] << 32) | (unsigned)stats[1]; dbg[4096 + (size_t)n * 12 + (size_t)e * 4 + 2] = ((long long)stats[2] << 32) | (unsigned)stats[3]; dbg[4096 + (size_t)n * 12 + (size_t)e * 4 + 3] = (long long)any | ((long long)stats[4] << 8) | ((long long)stats[5] << 36); }
}

/* The same answers, with the work where it is (measured on B200, tools/collide_tail.py: 93 % of the cars in a steady-state rollout only ever reach
 * the floor-box x TRACK-triangle loop, where a warp per car spends 17 k cycles on 5 rounds of at most 32 entries; 7 % have WALL triangles in the
 * hull's height range, 0.3 % reach the narrow phase): ONE THREAD PER CAR runs the floor test (car_collide, walls left out) and notes whether the
 * footprint holds wall triangles at hull height; the warp then takes those cars one after the other through the hull x WALL test with all
 * 32 lanes (car_collide_warp, floor left out).  Same predicates on the same data as k_collide, any-hit: identical flags. */
template <int LPC>      /* lanes per car of the floor test: the cells of the footprint are dealt to them (car_collide's cellPart / cellParts) */
__global__ void __launch_bounds__(PD_COLLIDE_BLOCK, PD_COLLIDE_MINBLOCKS) k_collide2(const __grid_constant__ PdCarParams P, const __grid_constant__ TrackDev T, const uint32_t* __restrict__ state, int layout, int n,
                                                              const int32_t* __restrict__ pending, int32_t* __restrict__ collOut,
                                                              int teleportMode, uint64_t seed, uint64_t idOffset, const uint32_t* __restrict__ episodeCtr) {
    const int lane = threadIdx.x & 31, sub = lane % LPC;
    const int e = (blockIdx.x * PD_COLLIDE_BLOCK + threadIdx.x) / LPC;
    const unsigned FULL = 0xffffffffu;
    __shared__ __align__(16) float hullS[(PD_COLLIDE_BLOCK / 32) * PD_HULLS_WORDS];
    Body C; C.fr.p = v3(0, 0, 0); C.fr.ax = v3(1, 0, 0); C.fr.ay = v3(0, 1, 0); C.fr.az = v3(0, 0, 1);
    bool hit = false, walls = false;
    if (e < n) {
        SVR sv = sv_env(layout, const_cast<uint32_t*>(state), (size_t)e);
        if (sv.i(PD_OFF_CAR + PD_CAR_o_physFrame) & 1) {
            load_body(sv, PD_BODY_CHASSIS, C);
            if (pending && pending[e]) {       /* reset inside the coming tick kernel, before its collision step: the pose the teleport will give (see k_collide) */
                float u = 0.0f;
                if (teleportMode == PD_TELEPORT_NEAREST) u = sv.f(PD_OFF_CAR + PD_CAR_o_trackLocation);
                else if (teleportMode == PD_TELEPORT_RANDOM) u = pd_uniform(seed, idOffset + (uint64_t)e, episodeCtr[e]);
                Quat q; teleport_chassis_pose(P, T, point_id_at_distance(T, u), C.fr.ax, C.fr.ay, C.fr.az, q, C.fr.p);
            }
            hit = car_collide(P, T, C, 0, 1, sub, LPC, 1, &walls);
        }
    }
    PD_UNROLL
    for (int off = 1; off < LPC; off <<= 1) { hit = __shfl_xor_sync(FULL, (int)hit, off) || hit; walls = __shfl_xor_sync(FULL, (int)walls, off) || walls; }
    if (hit) walls = false;
    unsigned m = __ballot_sync(FULL, walls && sub == 0);
    while (m) {
        const int k = __ffs(m) - 1; m &= m - 1;
        Body Cc;
        Cc.fr.p = v3(__shfl_sync(FULL, C.fr.p.x, k), __shfl_sync(FULL, C.fr.p.y, k), __shfl_sync(FULL, C.fr.p.z, k));
        Cc.fr.ax = v3(__shfl_sync(FULL, C.fr.ax.x, k), __shfl_sync(FULL, C.fr.ax.y, k), __shfl_sync(FULL, C.fr.ax.z, k));
        Cc.fr.ay = v3(__shfl_sync(FULL, C.fr.ay.x, k), __shfl_sync(FULL, C.fr.ay.y, k), __shfl_sync(FULL, C.fr.ay.z, k));
        Cc.fr.az = v3(__shfl_sync(FULL, C.fr.az.x, k), __shfl_sync(FULL, C.fr.az.y, k), __shfl_sync(FULL, C.fr.az.z, k));
        const bool wh = car_collide_warp<true, false>(P, T, Cc, lane, hullS + (threadIdx.x >> 5) * PD_HULLS_WORDS, nullptr);
        if (lane == k && wh) hit = true;
        __syncwarp(FULL);        /* the next car re-stages the hull tables of this warp's slot only after every lane is done with them */
    }
    if (e < n && sub == 0) collOut[e] = hit ? 1 : 0;
}

/* (k_collide2 as TWO launches -- the floor test alone at 64 registers / twice the resident warps, then a warp per flagged car for the walls -- was
 * measured slower: 78.9 vs 83.4 M car-ticks/s at 65536 envs, 31.2 vs 32.8 M at 16384.) */
/* Collision response: the contact joints of this odd frame (PhysicsEngineODE.cpp:230-236: the frame's contact group is emptied, then
 * refilled).  One warp per car, after k_collide on the same stream; cars on an even frame keep the joints of the previous frame,
 * cars without contact get an empty set, cars WITH contact (rare) run the generator of pd_contacts.h on the same start pose
 * k_collide tested. */
__global__ void __launch_bounds__(PD_COLLIDE_BLOCK) k_contacts(const __grid_constant__ PdCarParams P, const __grid_constant__ TrackDev T, const uint32_t* __restrict__ state, int layout, int n,
                                                               const int32_t* __restrict__ pending, const int32_t* __restrict__ collIn, float* __restrict__ contOut,
                                                               int teleportMode, uint64_t seed, uint64_t idOffset, const uint32_t* __restrict__ episodeCtr) {
    const int e = (blockIdx.x * PD_COLLIDE_BLOCK + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (e >= n) return;
    SVR sv = sv_env(layout, const_cast<uint32_t*>(state), (size_t)e);
    if (!(sv.i(PD_OFF_CAR + PD_CAR_o_physFrame) & 1)) return;
    float* co = contOut + (size_t)e * PD_CONTACT_WORDS;
    if (!collIn[e]) { if (lane == 0) reinterpret_cast<int*>(co)[0] = 0; return; }
    Body C; load_body(sv, PD_BODY_CHASSIS, C);
    if (pending && pending[e]) {
        float u = 0.0f;
        if (teleportMode == PD_TELEPORT_NEAREST) u = sv.f(PD_OFF_CAR + PD_CAR_o_trackLocation);
        else if (teleportMode == PD_TELEPORT_RANDOM) u = pd_uniform(seed, idOffset + (uint64_t)e, episodeCtr[e]);
        Quat q; teleport_chassis_pose(P, T, point_id_at_distance(T, u), C.fr.ax, C.fr.ay, C.fr.az, q, C.fr.p);
    }
    car_contacts_warp(P, T, C, lane, co);
}

/* initial record -> every env (both layouts) */
__global__ void k_broadcast(uint32_t* state, int layout, size_t nAlloc, const uint32_t* __restrict__ rec) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nAlloc) return;
    if (layout == PD_LAYOUT_RECORDS) { const int w = (int)(i % PD_STATE_STRIDE); state[i] = w < PD_STATE_WORDS ? rec[w] : 0u; }
    else state[i] = rec[(i / PD_TILE) % PD_STATE_WORDS];
}

/* Teleport (reset path): choose the spline point by mode (Car::teleportByMode, Car.cpp:1342-1358)
 * and re-seat the car there (Car::teleportToSpline); zeroAction: also write the reset's zero action and clear the
 * NaN guard (auto-reset inside pd_env_step, projectd_env.py:216-227). */
__global__ void __launch_bounds__(PD_BLOCK) k_teleport(const PdCarParams* __restrict__ P, TrackDev T, uint32_t* state, int layout, int n, const int32_t* __restrict__ mask,
                                                       int mode, const float* __restrict__ distNorm, uint64_t seed, uint64_t idOffset, uint32_t* __restrict__ episodeCtr,
                                                       double time, int zeroAction, int32_t* __restrict__ pending, PdEnvConfig cfg) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (mask && !mask[e]) return;
    if (pending) pending[e] = 0;      /* an explicit teleport supersedes a reset that was still due */
    SVR sv = sv_env(layout, state, (size_t)e);
    float u = 0.0f;
    if (distNorm) u = distNorm[e];
    else if (mode == PD_TELEPORT_NEAREST) u = sv.f(PD_OFF_CAR + PD_CAR_o_trackLocation);
    else if (mode == PD_TELEPORT_RANDOM) { u = pd_uniform(seed, idOffset + (uint64_t)e, episodeCtr[e]); episodeCtr[e]++; }
    car_teleport_to_point(*P, T, sv, point_id_at_distance(T, u), time);
    if (zeroAction) { sv.i(PD_OFF_CAR + PD_CAR_o_nanFlag, 0); env_apply_action(sv, 0.0f, 0.0f, cfg); }
}

__global__ void k_set_controls(uint32_t* state, int layout, int n, const float* __restrict__ ctl, const int8_t* __restrict__ gears, int smooth) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    SVR sv = sv_env(layout, state, (size_t)e);
    const int o = PD_OFF_CAR;
    sv.f(o + PD_CAR_o_ctlSteer, ctl[e * 5 + 0]); sv.f(o + PD_CAR_o_ctlClutch, ctl[e * 5 + 1]); sv.f(o + PD_CAR_o_ctlBrake, ctl[e * 5 + 2]);
    sv.f(o + PD_CAR_o_ctlHandBrake, ctl[e * 5 + 3]); sv.f(o + PD_CAR_o_ctlGas, ctl[e * 5 + 4]);
    sv.i(o + PD_CAR_o_ctlRequestedGear, gears ? (int)gears[e * 3 + 0] : -1);
    sv.i(o + PD_CAR_o_ctlGearUp, gears ? (int)gears[e * 3 + 1] : 0);
    sv.i(o + PD_CAR_o_ctlGearDn, gears ? (int)gears[e * 3 + 2] : 0);
    sv.i(o + PD_CAR_o_smoothSteer, smooth);
}

__global__ void k_set_actions(uint32_t* state, int layout, int n, const float* __restrict__ act, PdEnvConfig cfg) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    env_apply_action(sv_env(layout, state, (size_t)e), act[e * 2 + 0], act[e * 2 + 1], cfg);
}
__global__ void k_zero_counters(int n, const int32_t* __restrict__ mask, float* envReturn, int32_t* envLen) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || (mask && !mask[e])) return;
    envReturn[e] = 0.0f; envLen[e] = 0;
}

__global__ void k_observe(const uint32_t* state, int layout, int n, float* __restrict__ obs) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    SVR sv = sv_env(layout, const_cast<uint32_t*>(state), (size_t)e);
    float o[PD_OBS_DIM];
    car_observe(sv, o);
    for (int k = 0; k < PD_OBS_DIM; ++k) obs[(size_t)e * PD_OBS_DIM + k] = o[k];
}

__global__ void k_rewards(const uint32_t* state, int layout, int n, float* stepReward, float* totalReward, int32_t* flags) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    SVR sv = sv_env(layout, const_cast<uint32_t*>(state), (size_t)e);
    if (stepReward) stepReward[e] = sv.f(PD_OFF_CAR + PD_CAR_o_stepReward);
    if (totalReward) totalReward[e] = sv.f(PD_OFF_CAR + PD_CAR_o_totalReward);
    if (flags) flags[e] = (sv.i(PD_OFF_CAR + PD_CAR_o_collisionFlag) ? 1 : 0) | (sv.i(PD_OFF_CAR + PD_CAR_o_outOfTrackFlag) ? 2 : 0);
}

/* host record layout <-> device layout: words [PD_STATE_WORDS][n] on the host side (snapshot / restore) */
__global__ void k_pack(uint32_t* state, int layout, int n, uint32_t* __restrict__ soa, int toDevice) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * PD_STATE_WORDS) return;
    const int w = (int)(i / n); const size_t e = i % n;
    const size_t j = state_index(layout, w, e);
    if (toDevice) state[j] = soa[i]; else soa[i] = state[j];
}

__global__ void k_raycast(TrackDev T, int n, const float* __restrict__ rays, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = rays + (size_t)i * 7; float* q = out + (size_t)i * 8;
    const bool down = (p[3] == 0.0f && p[4] == -1.0f && p[5] == 0.0f);
    const RayHit r = down ? ray_cast_down(T, v3(p[0], p[1], p[2]), p[6]) : ray_cast(T, v3(p[0], p[1], p[2]), v3(p[3], p[4], p[5]), p[6]);
    q[0] = (float)r.hit; q[1] = r.pos.x; q[2] = r.pos.y; q[3] = r.pos.z; q[4] = r.normal.x; q[5] = r.normal.y; q[6] = r.normal.z; q[7] = (float)r.surface;
}

/* Track::computeFatPoints + computeSideLocation (Sim/Track.cpp:366-467), one thread per (spline point, side): the down ray that
 * drops the point onto the road, then either the two side rays at the stored half widths or the side TRACE -- rays from the
 * point's ray origin towards positions stepping outwards by `step`, each accepted only while the surface stays valid track of the
 * same category, height and grip continuous with the previous accepted hit and not in a bad sector; the first rejected hit ends
 * the walk.  Each walk is sequential by definition (it depends on its previous hit); the batch is parallel over points and sides. */
struct FatCfg { int traceSides; float offY, rayLen, sideMax, diffH, diffGrip, step; int nBad; uint32_t bad[8]; };
__global__ void k_fat_points(TrackDev T, int n, const float* __restrict__ slim, FatCfg cfg, float* __restrict__ fat /* [n][15] */) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = tid >> 1, side = tid & 1;
    if (i >= n) return;
    const float* sp = slim + (size_t)i * 5;
    float* f = fat + (size_t)i * 15;
    const V3 best0 = v3(sp[0], sp[1], sp[2]);
    const V3 rayOff = v3(0.0f, cfg.offY, 0.0f);
    const V3 rayStart = best0 + rayOff;
    const RayHit hit = ray_cast_down(T, rayStart, cfg.rayLen);
    if (!hit.hit) { if (side == 0) for (int k = 0; k < 15; ++k) f[k] = 0.0f; return; }
    V3 fwd = v3(0, 0, 0);
    if (i + 1 < n) fwd = norm(v3(slim[(size_t)(i + 1) * 5], slim[(size_t)(i + 1) * 5 + 1], slim[(size_t)(i + 1) * 5 + 2]) - best0);
    else if (i > 0) fwd = norm(best0 - v3(slim[(size_t)(i - 1) * 5], slim[(size_t)(i - 1) * 5 + 1], slim[(size_t)(i - 1) * 5 + 2]));
    const V3 leftDir = norm(cross(fwd, v3(0.0f, -1.0f, 0.0f)));
    const V3 dir = side == 0 ? leftDir : leftDir * -1.0f;
    V3 result;
    if (!cfg.traceSides) {
        result = hit.pos + dir * sp[3 + side];
        const RayHit h2 = ray_cast_down(T, result + rayOff, cfg.rayLen);
        if (h2.hit) result = h2.pos;
    } else {
        const PdSurface& s0 = T.surfaces[hit.surface];
        result = hit.pos; V3 prevHit = hit.pos; float prevGrip = s0.gripMod;
        const int numSteps = (int)(cfg.sideMax / cfg.step);
        for (int traceId = 1; traceId < numSteps; ++traceId) {
            const V3 rayEnd = hit.pos + dir * ((float)traceId * cfg.step);
            const V3 rayN = norm(rayEnd - rayStart);
            const RayHit h2 = ray_cast(T, rayStart, rayN, cfg.rayLen);
            if (!h2.hit) continue;
            const PdSurface& s2 = T.surfaces[h2.surface];
            bool bad = false;
            for (int q = 0; q < cfg.nBad; ++q) if (cfg.bad[q] == s2.sectorID) bad = true;
            if (s2.isValidTrack && s2.collisionCategory == s0.collisionCategory && fabsf(h2.pos.y - prevHit.y) < cfg.diffH && fabsf(s2.gripMod - prevGrip) < cfg.diffGrip && !bad) { result = h2.pos; prevHit = h2.pos; prevGrip = s2.gripMod; }
            else break;
        }
    }
    if (side == 0) { f[0] = hit.pos.x; f[1] = hit.pos.y; f[2] = hit.pos.z; f[3] = result.x; f[4] = result.y; f[5] = result.z; f[12] = fwd.x; f[13] = fwd.y; f[14] = fwd.z; }
    else { f[6] = result.x; f[7] = result.y; f[8] = result.z; }
}

__global__ void k_set_pressure(uint32_t* state, int layout, int n, int wheel, float value) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    state[state_index(layout, PD_OFF_TYRE(wheel) + PD_TYRE_o_pressureStatic, (size_t)e)] = f2u(value);
}

/* ------------------------------------------------------------------ host side ------------------------------------------------------------------ */
/* minimal DLPack ABI (dlpack.h v0.8) */
typedef struct { int32_t device_type; int32_t device_id; } PdDLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } PdDLDataType;
typedef struct { void* data; PdDLDevice device; int32_t ndim; PdDLDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset; } PdDLTensor;
typedef struct PdDLManagedTensor { PdDLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct PdDLManagedTensor*); } PdDLManagedTensor;

#pragma pack(push, 4)
struct PdCarStateOut {    /* Car/CarState.h:11-56 */
    int32_t carId, simId; float timestamp;
    float steer, clutch, brake, handBrake, gas; int8_t isShifterSupported, requestedGearIndex, gearUp, gearDn;
    int32_t collisionFlag, outOfTrackFlag, trackPointId; float lastTrackPointTimestamp, trackLocation, bodyVsTrack, velocityVsTrack;
    float engineRPM, speedMS; int32_t gear, gearGrinding;
    float bodyMatrix[16], bodyPos[3], bodyEuler[3], accG[3], velocity[3], localVelocity[3], angularVelocity[3], localAngularVelocity[3];
    float hubMatrix[4][16], tyreContacts[4][3], tyreLoad[4], tyreAngularSpeed[4], tyreSlipRatio[4], tyreNdSlip[4];
    float probes[10], lookAhead[5], stepReward, totalReward;
};
#pragma pack(pop)
static_assert(sizeof(PdCarStateOut) == 664, "CarState is 664 bytes");

/* batches up to PD_QUAD_MAX_ENVS: four lanes per car (latency-bound regime, more warps per car);
 * larger batches: one thread per car (throughput regime, no redundant scalar work) */
#define PD_QUAD_MAX_ENVS 20480     /* measured crossover on B200 (round 2, k_tick_quad<8> against k_tick + k_collide2, M car-ticks/s): 12288 envs 37.2 / 26.7, 16384: 41.9 / 33.0, 24576: 43.8 / 47.8, 32768: 48.8 / 56.4 */

struct pd_batch {
    int n = 0, device = 0;
    cudaStream_t stream = nullptr, ownStream = nullptr;
    PdEnvConfig envCfg{0.1f, 1.0f, 50.0f, 50.0f, 50.0f, -200.0f, 5.0f, 1, 1, 1, 1, 0.0f, -1};   /* projectd_env.py:27-53 */
    pdh::CarModel car; pdh::TrackModel track;
    PdCarParams* dP = nullptr; bool paramsDirty = true;
    TrackDev dev{};
    std::vector<void*> allocs;
    uint32_t* dState = nullptr;
    float* dObs = nullptr; float* dCtl = nullptr; int8_t* dGears = nullptr; float* dAct = nullptr;
    int32_t* dMask = nullptr; int32_t* dPoints = nullptr; float* dDist = nullptr; uint32_t* dEpisodeCtr = nullptr;
    float* dReward = nullptr; float* dTotal = nullptr; int32_t* dFlags = nullptr; int32_t* dDone = nullptr;
    float* dEnvReturn = nullptr; int32_t* dEnvLen = nullptr; double* dStats = nullptr;
    long long* dClk = nullptr; int nClk = 0;
    bool serialWide = false;          /* thread-per-car kernel with the 255-register budget (batches that fit one wave at 4 blocks per SM; env PD_SERIAL_WIDE overrides) */
    int serialSmemPad = 0;            /* tuning knob (env PD_SERIAL_SMEM_PAD, bytes): unused dynamic shared memory per block of k_tick, caps the resident blocks per SM */
    int collideLpc = 4;               /* env PD_COLLIDE_LPC: lanes per car of k_collide2's floor test (1 / 4 / 8 / 16); measured at 65536 envs: 79.4 / 81.1 / 80.2 / 78.2 M car-ticks/s (warp per car, k_collide: 75.3 M) */
    bool debugSkipCollision = false;  /* env PD_DEBUG_SKIP_COLLISION=1: MEASUREMENT ONLY -- no collision test at all (wrong flags), to see what the detection costs a tick */
    int bvhOnDevice = 0, bvhDepth = 0; /* the ray caster's tree was built by pd_lbvh.h (and its depth) */
    bool collideV1 = false;           /* env PD_COLLIDE_V1=1: k_collide (a warp per car) instead of k_collide2 (a thread per car for the floor, a warp for the walls) */
    bool inlineCollide = false;       /* thread-per-car kernel: test collisions inside the tick (env PD_SERIAL_INLINE_COLLIDE=1) instead of k_collide ahead of it */
    bool collWarp = true;             /* quad kernel: collision warp inside the tick kernel (env PD_COLL_WARP=0: k_collide ahead of it instead) */
    bool zeroCopy = true; const void* zcKey[4] = {nullptr, nullptr, nullptr, nullptr}; void* zcDev[4] = {nullptr, nullptr, nullptr, nullptr};
    int32_t* dColl = nullptr;         /* k_collide's answers for the coming tick */
    float* dContacts = nullptr;       /* [n][PD_CONTACT_WORDS] contact joints alive per env (collision response) */
    bool response = true;             /* A14 response: contact joints from collisions enter the solve (pd_set_collision_response) */
    long long frameKnown = 0;         /* physics frame shared by all envs, or -1 when states were set individually (then k_collide runs every tick) */
    int32_t* dPending = nullptr; int autoreset = PD_AUTORESET_SAME_STEP;
    int resetMode = PD_TELEPORT_START;   /* ProjectDEnv.teleport_mode: the last pd_teleport_mode() mode, also used by the automatic resets */
    double time = 0, lastDt = 0;
    uint64_t seed = 0, idOffset = 0, launches = 0;
    int layout = PD_LAYOUT_TILED;     /* PD_LAYOUT_RECORDS when the 4-lanes-per-car kernel owns the batch */
    int quadCpw = 8;                  /* cars per warp of the quad kernel (8 / 4 / 2), chosen at creation from the batch size */
    int quadMax = PD_QUAD_MAX_ENVS;   /* kernel dispatch threshold; env PD_QUAD_MAX_ENVS overrides (tuning / profiling) */
    std::string err;
    int64_t dlShape[2] = {0, 0};
};

static std::string g_createError;

/* every ABI entry that touches the device makes the batch's device current first (the caller's current device may have changed) */
#define ENTER(b) do { if (cudaSetDevice((b)->device) != cudaSuccess) { (b)->err = "cudaSetDevice failed"; return PD_ERR_CUDA; } } while (0)
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { b->err = std::string(#call) + ": " + cudaGetErrorString(e_); return PD_ERR_CUDA; } } while (0)

template <class T> static int dalloc(pd_batch* b, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 4);
    if (e != cudaSuccess) { b->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return PD_ERR_CUDA; }
    b->allocs.push_back(q); *p = (T*)q; return PD_OK;
}
template <class T> static int upload(pd_batch* b, const T** p, const std::vector<T>& v) {
    T* q = nullptr; int rc = dalloc(b, &q, v.size()); if (rc) return rc;
    if (!v.empty()) CK(cudaMemcpyAsync(q, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, b->stream));
    *p = q; return PD_OK;
}
static inline int grid(int n, int block) { return (n + block - 1) / block; }
/* small batches: one warp per block so that every SM gets work (148 SMs) */
static inline int tick_block(int) { return PD_QBLOCK; }

static int sync_params(pd_batch* b) {
    if (!b->paramsDirty) return PD_OK;
    CK(cudaMemcpyAsync(b->dP, &b->car.P, sizeof(PdCarParams), cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream));   /* the host copy may change again right after */
    b->paramsDirty = false; return PD_OK;
}

/* Track::computeFatPoints on the device: needs only the ray structures of the track (BVH, triangles, surfaces, column grid), which
 * are uploaded into temporaries here; fills track.fat and builds the point grids (pdh::finish_track_points). */
static int fat_points_on_device(pdh::TrackModel& track, cudaStream_t stream, std::string& err) {
    const int n = (int)(track.slim.size() / 5);
    if (n <= 0) { err = "no spline points"; return PD_ERR_IO; }
    TrackDev T{}; std::vector<void*> tmp;
    auto up = [&](const void* src, size_t bytes) -> void* { void* q = nullptr; if (cudaMalloc(&q, bytes ? bytes : 4) != cudaSuccess) return nullptr; tmp.push_back(q); if (bytes && src) cudaMemcpyAsync(q, src, bytes, cudaMemcpyHostToDevice, stream); return q; };
    T.nodes = (const BvhNode*)up(track.nodes.data(), track.nodes.size() * sizeof(pdh::BvhNodeH));
    T.tris = (const float*)up(track.tris.data(), track.tris.size() * 4);
    T.triSurf = (const int32_t*)up(track.triSurf.data(), track.triSurf.size() * 4);
    T.surfaces = (const PdSurface*)up(track.surfaces.data(), track.surfaces.size() * sizeof(PdSurface));
    T.colStart = (const int32_t*)up(track.colStart.data(), track.colStart.size() * 4);
    T.colItems = (const int32_t*)up(track.colItems.data(), track.colItems.size() * 4);
    T.colGrid = track.colGrid; T.info = track.info;
    float* dSlim = (float*)up(track.slim.data(), track.slim.size() * 4);
    float* dFat = (float*)up(nullptr, (size_t)n * 15 * 4);
    int rc = PD_OK;
    if (!T.nodes || !T.tris || !T.triSurf || !T.surfaces || !T.colStart || !T.colItems || !dSlim || !dFat) { err = "cudaMalloc failed (fat points)"; rc = PD_ERR_CUDA; }
    if (rc == PD_OK) {      /* the traces through the device-built tree where pd_create would use it (same rule: PD_DEVICE_BVH, or >= 400 000 triangles) */
        const char* q = getenv("PD_DEVICE_BVH");
        const int nt = track.info.nTris;
        if ((q ? atoi(q) != 0 : nt >= 400000) && nt >= 2) {
            float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            for (size_t i = 0; i < track.triRaw.size(); ++i) { const int k = (int)(i % 3); lo[k] = std::min(lo[k], track.triRaw[i]); hi[k] = std::max(hi[k], track.triRaw[i]); }
            const float* dRaw = (const float*)up(track.triRaw.data(), track.triRaw.size() * 4);
            BvhNode* nodes = nullptr; int count = 0, depth = 0;
            const char* e = dRaw ? lbvh_build_on_device(dRaw, nt, lo, hi, stream, &nodes, &count, &depth) : "cudaMalloc failed";
            if (!e) { tmp.push_back(nodes); T.nodes = nodes; T.info.nNodes = count; }
            else if (q) { err = std::string("device BVH build failed: ") + e; rc = PD_ERR_CUDA; }
        }
    }
    if (rc == PD_OK) {
        FatCfg cfg{}; const pdh::TraceConfig& tc = track.trace;
        cfg.traceSides = tc.traceSides; cfg.offY = tc.rayOffsetY; cfg.rayLen = tc.rayLength; cfg.sideMax = tc.sideMax; cfg.diffH = tc.diffHeightMax; cfg.diffGrip = tc.diffGripMax; cfg.step = tc.step;
        cfg.nBad = tc.nBadSectors; for (int q = 0; q < 8; ++q) cfg.bad[q] = tc.badSectors[q];
        cudaMemsetAsync(dFat, 0, (size_t)n * 15 * 4, stream);
        k_fat_points<<<(2 * n + 63) / 64, 64, 0, stream>>>(T, n, dSlim, cfg, dFat);
        track.fat.resize((size_t)n);
        if (cudaMemcpyAsync(track.fat.data(), dFat, (size_t)n * 15 * 4, cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) { err = std::string("k_fat_points: ") + cudaGetErrorString(cudaGetLastError()); rc = PD_ERR_CUDA; }
    }
    for (void* q : tmp) cudaFree(q);
    if (rc != PD_OK) return rc;
    for (PdFatPoint& p : track.fat) for (int k = 0; k < 3; ++k) p.center[k] = (p.left[k] + p.right[k]) * 0.5f;     /* fat.center = (left + right) * 0.5 (Track.cpp:430) */
    track.needFat = false;
    try { pdh::finish_track_points(track, track.closedLoop, track.hashCellSize); } catch (const std::exception& ex) { err = ex.what(); return PD_ERR_IO; }
    return PD_OK;
}

static int finish_create(pd_batch* b, int n_envs, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) { b->err = "no CUDA device available (this library has no CPU path)"; return PD_ERR_CUDA; }
    if (device < 0 || device >= count) { b->err = "bad device ordinal"; return PD_ERR_ARG; }
    CK(cudaSetDevice(device));
    b->n = n_envs; b->device = device;
    if (const char* q = getenv("PD_QUAD_MAX_ENVS")) b->quadMax = atoi(q);
    if (const char* q = getenv("PD_E2E_ZEROCOPY")) b->zeroCopy = atoi(q) != 0;
    if (const char* q = getenv("PD_COLL_WARP")) b->collWarp = atoi(q) != 0;
    if (const char* q = getenv("PD_SERIAL_INLINE_COLLIDE")) b->inlineCollide = atoi(q) != 0;
    if (const char* q = getenv("PD_COLLIDE_V1")) b->collideV1 = atoi(q) != 0;
    if (const char* q = getenv("PD_DEBUG_SKIP_COLLISION")) b->debugSkipCollision = atoi(q) != 0;
    if (const char* q = getenv("PD_COLLIDE_LPC")) b->collideLpc = atoi(q);
    if (const char* q = getenv("PD_SERIAL_SMEM_PAD")) { b->serialSmemPad = atoi(q); cudaFuncSetAttribute(k_tick<PD_TOPO_STRUT_AXLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, b->serialSmemPad); cudaFuncSetAttribute(k_tick<PD_TOPO_STRUT_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, b->serialSmemPad); cudaFuncSetAttribute(k_tick<PD_TOPO_DW_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, b->serialSmemPad); }
    CK(cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking)); b->ownStream = b->stream;
    CK(cudaFuncSetAttribute(k_tick_quad<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(8)));
    CK(cudaFuncSetAttribute(k_tick_quad<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(4)));
    CK(cudaFuncSetAttribute(k_tick_quad<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(2)));
    CK(cudaFuncSetAttribute((k_tick_quad<8, PD_TOPO_STRUT_DW>), cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(8)));
    CK(cudaFuncSetAttribute((k_tick_quad<4, PD_TOPO_STRUT_DW>), cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(4)));
    CK(cudaFuncSetAttribute((k_tick_quad<8, PD_TOPO_DW_DW>), cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(8)));
    CK(cudaFuncSetAttribute((k_tick_quad<4, PD_TOPO_DW_DW>), cudaFuncAttributeMaxDynamicSharedMemorySize, PD_QUAD_SMEM_BYTES_(4)));
    /* cars per warp of the quad kernel: as few as still fit the batch into ONE wave of resident blocks
       (168 registers x 64 threads -> 6 blocks per SM; shared memory allows 2 / 4 / 8 blocks for 8 / 4 / 2 cars per warp) */
    { cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, device)); const int sms = prop.multiProcessorCount;
      b->quadCpw = (n_envs <= sms * 6 * 4) ? 2 : (n_envs <= sms * 4 * 8) ? 4 : 8;
      if (const char* q = getenv("PD_QUAD_CPW")) { const int v = atoi(q); if (v == 2 || v == 4 || v == 8) b->quadCpw = v; }
      b->serialWide = n_envs <= sms * 4 * PD_BLOCK;
      if (const char* q = getenv("PD_SERIAL_WIDE")) b->serialWide = atoi(q) != 0; }
    b->layout = (n_envs <= b->quadMax) ? PD_LAYOUT_RECORDS : PD_LAYOUT_TILED;
    if (b->car.P.topology != PD_TOPO_STRUT_AXLE && b->quadCpw == 2) b->quadCpw = 4;      /* double-wishbone cars: the 4-lanes-per-car kernel is instantiated for 4 and 8 cars per warp */
    int rc;
    if (b->track.needFat) { b->launches++; if ((rc = fat_points_on_device(b->track, b->stream, b->err))) return rc; }      /* no usable spline.cache: Track::computeFatPoints, on the GPU */
    if ((rc = dalloc(b, &b->dP, 1))) return rc;
    { const std::vector<pd::BvhNode>& dummy = *reinterpret_cast<const std::vector<pd::BvhNode>*>(&b->track.nodes); if ((rc = upload(b, &b->dev.nodes, dummy))) return rc; }
    if ((rc = upload(b, &b->dev.tris, b->track.tris))) return rc;
    if ((rc = upload(b, &b->dev.triSurf, b->track.triSurf))) return rc;
    if ((rc = upload(b, &b->dev.surfaces, b->track.surfaces))) return rc;
    if ((rc = upload(b, &b->dev.fat, b->track.fat))) return rc;
    if ((rc = upload(b, &b->dev.splineXYZ, b->track.splineXYZ))) return rc;
    if ((rc = upload(b, &b->dev.splineDist, b->track.splineDist))) return rc;
    if ((rc = upload(b, &b->dev.segStart, b->track.segStart))) return rc;
    if ((rc = upload(b, &b->dev.segItems, b->track.segItems))) return rc;
    if ((rc = upload(b, &b->dev.ptStart, b->track.ptStart))) return rc;
    if ((rc = upload(b, &b->dev.ptItems, b->track.ptItems))) return rc;
    if ((rc = upload(b, &b->dev.segRec, b->track.segRec))) return rc;
    if ((rc = upload(b, &b->dev.ptRec, b->track.ptRec))) return rc;
    b->dev.grid = b->track.grid; b->dev.segGrid = b->track.segGrid;
    if ((rc = upload(b, &b->dev.colStart, b->track.colStart))) return rc;
    if ((rc = upload(b, &b->dev.colItems, b->track.colItems))) return rc;
    b->dev.colGrid = b->track.colGrid;
    if ((rc = upload(b, &b->dev.triRaw, b->track.triRaw))) return rc;
    if ((rc = upload(b, &b->dev.collStart, b->track.collStart))) return rc;
    if ((rc = upload(b, &b->dev.collItems, b->track.collItems))) return rc;
    if ((rc = upload(b, &b->dev.collCell, b->track.collCell))) return rc;
    if ((rc = upload(b, &b->dev.collRec, b->track.collRec))) return rc;
    if ((rc = upload(b, &b->dev.collPlane, b->track.collPlane))) return rc;
    b->dev.collGrid = b->track.collGrid;
    {   /* the hull's tables in car_collide_warp's layout */
        std::vector<float> ht(PD_HULLS_WORDS, 0.0f); const PdCarParams& Pc = b->car.P;
        for (int i = 0; i < PD_MAX_COLLIDER_TRIS; ++i) {
            for (int q = 0; q < 4; ++q) ht[PD_HULLS_SPHERE + i * 4 + q] = Pc.colliderTriSphere[i][q];
            for (int q = 0; q < 6; ++q) ht[PD_HULLS_BOUNDS + i * 6 + q] = Pc.colliderTriBounds[i][q];
            const int32_t tri = (int)Pc.colliderTris[i][0] | ((int)Pc.colliderTris[i][1] << 8) | ((int)Pc.colliderTris[i][2] << 16);
            memcpy(&ht[PD_HULLS_TRIS + i], &tri, 4);
        }
        for (int i = 0; i < PD_MAX_COLLIDER_VERTS; ++i) for (int q = 0; q < 3; ++q) ht[PD_HULLS_VERTS + i * 3 + q] = Pc.colliderVerts[i][q];
        if ((rc = upload(b, &b->dev.hullTables, ht))) return rc;
    }
    b->dev.info = b->track.info;
    {   /* the ray caster's tree built ON THE DEVICE (pd_lbvh.h) for very large tracks (>= 400 000 triangles: BASELINE configs[3]) or on request
           (PD_DEVICE_BVH=1; =0 keeps the host-built tree).  Closest-hit rays do not depend on which tree they walk. */
        const char* q = getenv("PD_DEVICE_BVH");
        const bool want = q ? atoi(q) != 0 : b->track.info.nTris >= 400000;
        const int nt = b->track.info.nTris;
        if (want && nt >= 2) {
            float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            for (size_t i = 0; i < b->track.triRaw.size(); ++i) { const int k = (int)(i % 3); lo[k] = std::min(lo[k], b->track.triRaw[i]); hi[k] = std::max(hi[k], b->track.triRaw[i]); }
            pd::BvhNode* nodes = nullptr; int count = 0, depth = 0;
            const char* e = pd::lbvh_build_on_device(b->dev.triRaw, nt, lo, hi, b->stream, &nodes, &count, &depth);
            b->launches += 8;
            if (!e) { b->allocs.push_back(nodes); b->dev.nodes = nodes; b->dev.info.nNodes = count; b->bvhOnDevice = 1; b->bvhDepth = depth; }
            else if (q) { b->err = std::string("device BVH build failed: ") + e; return PD_ERR_CUDA; }      /* an explicit request does not fall back silently */
        }
    }
    const size_t n = (size_t)n_envs;
    const size_t nAlloc = state_alloc_words(b->layout, n);
    if ((rc = dalloc(b, &b->dState, nAlloc))) return rc;
    if ((rc = dalloc(b, &b->dObs, n * PD_OBS_DIM))) return rc;
    if ((rc = dalloc(b, &b->dCtl, n * 5))) return rc;
    if ((rc = dalloc(b, &b->dGears, n * 3))) return rc;
    if ((rc = dalloc(b, &b->dAct, n * 2))) return rc;
    if ((rc = dalloc(b, &b->dMask, n))) return rc;
    if ((rc = dalloc(b, &b->dPoints, n))) return rc;
    if ((rc = dalloc(b, &b->dDist, n))) return rc;
    if ((rc = dalloc(b, &b->dEpisodeCtr, n))) return rc;
    if ((rc = dalloc(b, &b->dReward, n))) return rc;
    if ((rc = dalloc(b, &b->dTotal, n))) return rc;
    if ((rc = dalloc(b, &b->dFlags, n))) return rc;
    if ((rc = dalloc(b, &b->dDone, n))) return rc;
    if ((rc = dalloc(b, &b->dEnvReturn, n))) return rc;
    if ((rc = dalloc(b, &b->dEnvLen, n))) return rc;
    if ((rc = dalloc(b, &b->dStats, 8))) return rc;
    if ((rc = dalloc(b, &b->dPending, n))) return rc;
    if ((rc = dalloc(b, &b->dColl, n))) return rc;
    if ((rc = dalloc(b, &b->dContacts, n * PD_CONTACT_WORDS))) return rc;
    CK(cudaMemsetAsync(b->dContacts, 0, n * PD_CONTACT_WORDS * 4, b->stream));
    CK(cudaMemsetAsync(b->dPending, 0, n * 4, b->stream));
    if (getenv("PD_DEBUG_CLOCKS")) { b->nClk = (int)(n / 4 + 64);
#if defined(PD_PHASE_CLOCKS)
        b->nClk = 4096 + (int)n * 16;
#endif
        if ((rc = dalloc(b, &b->dClk, (size_t)b->nClk))) return rc; CK(cudaMemsetAsync(b->dClk, 0, (size_t)b->nClk * 8, b->stream)); }
    CK(cudaMemsetAsync(b->dEpisodeCtr, 0, n * 4, b->stream));
    CK(cudaMemsetAsync(b->dEnvReturn, 0, n * 4, b->stream));
    CK(cudaMemsetAsync(b->dEnvLen, 0, n * 4, b->stream));
    CK(cudaMemsetAsync(b->dStats, 0, 8 * 8, b->stream));
    CK(cudaMemsetAsync(b->dObs, 0, n * PD_OBS_DIM * 4, b->stream));
    /* initial record: built once with the same device functions compiled for the host, then broadcast */
    std::vector<uint32_t> rec(PD_STATE_WORDS, 0);
    { SVFlat sv = sv_flat(rec.data()); car_init_state(b->car.P, sv); }
    uint32_t* dRec = nullptr; if ((rc = dalloc(b, &dRec, PD_STATE_WORDS))) return rc;
    CK(cudaMemcpyAsync(dRec, rec.data(), PD_STATE_WORDS * 4, cudaMemcpyHostToDevice, b->stream));
    k_broadcast<<<(unsigned)((nAlloc + 255) / 256), 256, 0, b->stream>>>(b->dState, b->layout, nAlloc, dRec); b->launches++;
    CK(cudaGetLastError());
    if ((rc = sync_params(b))) return rc;
    CK(cudaStreamSynchronize(b->stream));
    return PD_OK;
}

static void launch_tick(pd_batch* b, float dt, const int32_t* mask, const EnvIO& io_in) {
    EnvIO io = io_in; io.clk = mask ? nullptr : b->dClk; io.cfg = b->envCfg;
    bool collWarp = false;
    if (!mask) {
        /* collision detection (odd physics frames): the quad kernel brings its own collision warp per block; the thread-per-car
           kernel is preceded by k_collide (a warp per car) on the same stream */
        const bool oddPossible = (b->frameKnown < 0 || (b->frameKnown & 1)) && !b->debugSkipCollision;
        if (b->debugSkipCollision) { cudaMemsetAsync(b->dColl, 0, (size_t)b->n * 4, b->stream); io.collIn = b->dColl; }       /* "no contact" for every car: measurement only */
        /* with the collision response on, k_collide also GENERATES the contact joints of cars that touch something (it knows the
           start pose, resets included); the tick kernels pick the live joints up at the solve */
        if (oddPossible && b->layout == PD_LAYOUT_RECORDS && b->collWarp && !b->response) collWarp = true;
        else if (oddPossible && b->layout == PD_LAYOUT_TILED && b->inlineCollide && !b->response) { /* experiment: every thread tests its own car inside k_tick */ }
        else if (oddPossible) {
            if (b->collideV1 || b->dClk)   /* the warp-per-car form (round 1; keeps the per-car clocks of the profiling tools) */
                k_collide<<<grid(b->n, PD_COLLIDE_BLOCK / 32), PD_COLLIDE_BLOCK, 0, b->stream>>>(b->car.P, b->dev, b->dState, b->layout, b->n, io.pending, b->dColl, (b->dClk && b->nClk >= 4096 + b->n * 16) ? b->dClk : nullptr,
                                                                                                          io.teleportMode, io.seed, io.idOffset, io.episodeCtr);
            else
                switch (b->collideLpc) {
                case 1: k_collide2<1><<<grid(b->n, PD_COLLIDE_BLOCK), PD_COLLIDE_BLOCK, 0, b->stream>>>(b->car.P, b->dev, b->dState, b->layout, b->n, io.pending, b->dColl, io.teleportMode, io.seed, io.idOffset, io.episodeCtr); break;
                default: k_collide2<4><<<grid(b->n, PD_COLLIDE_BLOCK / 4), PD_COLLIDE_BLOCK, 0, b->stream>>>(b->car.P, b->dev, b->dState, b->layout, b->n, io.pending, b->dColl, io.teleportMode, io.seed, io.idOffset, io.episodeCtr); break;
                case 16: k_collide2<16><<<grid(b->n, PD_COLLIDE_BLOCK / 16), PD_COLLIDE_BLOCK, 0, b->stream>>>(b->car.P, b->dev, b->dState, b->layout, b->n, io.pending, b->dColl, io.teleportMode, io.seed, io.idOffset, io.episodeCtr); break;
                case 8: k_collide2<8><<<grid(b->n, PD_COLLIDE_BLOCK / 8), PD_COLLIDE_BLOCK, 0, b->stream>>>(b->car.P, b->dev, b->dState, b->layout, b->n, io.pending, b->dColl, io.teleportMode, io.seed, io.idOffset, io.episodeCtr); break;
                }
            b->launches++;
            io.collIn = b->dColl;
            if (b->response) { k_contacts<<<grid(b->n, PD_COLLIDE_BLOCK / 32), PD_COLLIDE_BLOCK, 0, b->stream>>>(b->car.P, b->dev, b->dState, b->layout, b->n, io.pending, b->dColl, b->dContacts,
                                                                                                                        io.teleportMode, io.seed, io.idOffset, io.episodeCtr); b->launches++; }
        }
        if (b->response) io.contacts = b->dContacts;
        if (b->frameKnown >= 0) b->frameKnown++;
    } else { b->frameKnown = -1; if (b->response) io.contacts = b->dContacts; }          /* a masked tick advances only some envs: frames are no longer in lock step */
    if (b->layout == PD_LAYOUT_RECORDS) {
        const int threads = collWarp ? PD_QBLOCK + 32 : PD_QBLOCK;
#define PD_LAUNCH_QUAD(CPW_, ...) k_tick_quad<CPW_, ##__VA_ARGS__><<<grid(b->n, 2 * CPW_), threads, PD_QUAD_SMEM_BYTES_(CPW_), b->stream>>>(b->car.P, b->dev, b->dState, b->n, dt, b->time, mask, io)
        const int topo = b->car.P.topology;
        if (topo == PD_TOPO_STRUT_DW) { if (b->quadCpw == 4) PD_LAUNCH_QUAD(4, PD_TOPO_STRUT_DW); else PD_LAUNCH_QUAD(8, PD_TOPO_STRUT_DW); }
        else if (topo == PD_TOPO_DW_DW) { if (b->quadCpw == 4) PD_LAUNCH_QUAD(4, PD_TOPO_DW_DW); else PD_LAUNCH_QUAD(8, PD_TOPO_DW_DW); }
        else switch (b->quadCpw) {
        case 2: PD_LAUNCH_QUAD(2); break;
        case 4: PD_LAUNCH_QUAD(4); break;
        default: PD_LAUNCH_QUAD(8); break;
        }
    } else
        switch (b->car.P.topology) {
        case PD_TOPO_STRUT_DW: k_tick<PD_TOPO_STRUT_DW><<<grid(b->n, PD_BLOCK), PD_BLOCK, b->serialSmemPad, b->stream>>>(b->car.P, b->dev, b->dState, b->n, dt, b->time, mask, io); break;
        case PD_TOPO_DW_DW: k_tick<PD_TOPO_DW_DW><<<grid(b->n, PD_BLOCK), PD_BLOCK, b->serialSmemPad, b->stream>>>(b->car.P, b->dev, b->dState, b->n, dt, b->time, mask, io); break;
        default:
            if (b->serialWide) k_tick<PD_TOPO_STRUT_AXLE, 4><<<grid(b->n, PD_BLOCK), PD_BLOCK, b->serialSmemPad, b->stream>>>(b->car.P, b->dev, b->dState, b->n, dt, b->time, mask, io);
            else k_tick<PD_TOPO_STRUT_AXLE><<<grid(b->n, PD_BLOCK), PD_BLOCK, b->serialSmemPad, b->stream>>>(b->car.P, b->dev, b->dState, b->n, dt, b->time, mask, io);
            break;
        }
    b->launches++;
}

extern "C" {

int pd_create(const char* base_path, const char* track_name, const char* car_model, int n_envs, int device, pd_batch** out) {
    if (!out || !base_path || !track_name || !car_model || n_envs <= 0) { g_createError = "pd_create: bad argument"; return PD_ERR_ARG; }
    *out = nullptr;
    pd_batch* b = new pd_batch();
    int rc = PD_OK;
    try { pdh::load_car(base_path, car_model, b->car); pdh::load_track(base_path, track_name, b->track); }
    catch (const std::exception& ex) { b->err = ex.what(); rc = PD_ERR_IO; }
    if (rc == PD_OK) rc = finish_create(b, n_envs, device);
    if (rc != PD_OK) { g_createError = b->err; pd_destroy(b); return rc; }
    *out = b; return PD_OK;
}

int pd_compute_fat_points(const char* base_path, const char* track_name, int device, float* out, int cap_points) {
    if (!base_path || !track_name || !out || cap_points < 0) { g_createError = "pd_compute_fat_points: bad argument"; return PD_ERR_ARG; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) { g_createError = "no CUDA device available (this library has no CPU path)"; return PD_ERR_CUDA; }
    if (device < 0 || device >= count || cudaSetDevice(device) != cudaSuccess) { g_createError = "bad device ordinal"; return PD_ERR_ARG; }
    pdh::TrackModel track;
    try { pdh::load_track(base_path, track_name, track, true); } catch (const std::exception& ex) { g_createError = ex.what(); return PD_ERR_IO; }
    cudaStream_t st; if (cudaStreamCreate(&st) != cudaSuccess) { g_createError = "cudaStreamCreate failed"; return PD_ERR_CUDA; }
    std::string err; const int rc = fat_points_on_device(track, st, err);
    cudaStreamDestroy(st);
    if (rc != PD_OK) { g_createError = err; return rc; }
    const int n = (int)track.fat.size();
    memcpy(out, track.fat.data(), (size_t)(n < cap_points ? n : cap_points) * sizeof(PdFatPoint));
    return n;
}

int pd_create_synthetic(const char* base_path, const char* car_model, int target_tris, float length_m, int n_envs, int device, pd_batch** out) {
    if (!out || !base_path || !car_model || n_envs <= 0 || target_tris <= 0) { g_createError = "pd_create_synthetic: bad argument"; return PD_ERR_ARG; }
    *out = nullptr;
    pd_batch* b = new pd_batch();
    int rc = PD_OK;
    try { pdh::load_car(base_path, car_model, b->car); pdh::make_synthetic_track(target_tris, length_m, b->track); }
    catch (const std::exception& ex) { b->err = ex.what(); rc = PD_ERR_IO; }
    if (rc == PD_OK) rc = finish_create(b, n_envs, device);
    if (rc != PD_OK) { g_createError = b->err; pd_destroy(b); return rc; }
    *out = b; return PD_OK;
}

void pd_destroy(pd_batch* b) {
    if (!b) return;
    cudaSetDevice(b->device);
    if (b->stream) cudaStreamSynchronize(b->stream);
    for (void* p : b->allocs) cudaFree(p);
    if (b->ownStream) cudaStreamDestroy(b->ownStream);
    delete b;
}
const char* pd_last_error(const pd_batch* b) { return b ? b->err.c_str() : g_createError.c_str(); }
int pd_num_envs(const pd_batch* b) { return b ? b->n : 0; }
int pd_state_words(void) { return PD_STATE_WORDS; }
int pd_obs_dim(void) { return PD_OBS_DIM; }
int pd_car_state_bytes(void) { return (int)sizeof(PdCarStateOut); }

int pd_set_assists(pd_batch* b, int ac, int as, int ab) {
    if (!b) return PD_ERR_ARG;
    PdAssists& A = b->car.P.assists;
    A.acUseAutoOnStart = ac != 0; A.acUseAutoOnChange = ac != 0; A.asIsActive = as != 0; A.blipIsActive = ab != 0;
    b->paramsDirty = true; return PD_OK;
}
static int tune_common(pd_batch* b, const char* name, float value, bool raw) {
    if (!b || !name) return PD_ERR_ARG;
    ENTER(b);
    try { if (raw) b->car.setRawTune(name, value); else b->car.setTune(name, value); } catch (const std::exception& ex) { b->err = ex.what(); return PD_ERR_UNSUPPORTED; }
    b->paramsDirty = true;
    static const char* kP[4] = {"PRESSURE_LF", "PRESSURE_RF", "PRESSURE_LR", "PRESSURE_RR"};
    for (int w = 0; w < 4; ++w) if (!strcmp(name, kP[w])) {   /* the tune writes Tyre::status.pressureStatic of every car */
        k_set_pressure<<<grid(b->n, 256), 256, 0, b->stream>>>(b->dState, b->layout, b->n, w, b->car.P.tyre[w].pressureStaticDefault); b->launches++;
        CK(cudaGetLastError());
    }
    return PD_OK;
}
int pd_set_tune(pd_batch* b, const char* name, float value) { return tune_common(b, name, value, false); }
int pd_set_raw_tune(pd_batch* b, const char* name, float value) { return tune_common(b, name, value, true); }
int pd_set_env_config(pd_batch* b, const PdEnvConfig* cfg) {
    if (!b || !cfg) return PD_ERR_ARG;
    if (!(cfg->max_gas >= cfg->min_gas) || cfg->stuck_timeout < 0.0f) { b->err = "pd_set_env_config: bad gas range / stuck timeout"; return PD_ERR_ARG; }
    b->envCfg = *cfg; return PD_OK;
}
int pd_get_env_config(const pd_batch* b, PdEnvConfig* out) { if (!b || !out) return PD_ERR_ARG; *out = b->envCfg; return PD_OK; }
int pd_env_reset_counters(pd_batch* b, const uint8_t* mask) {
    if (!b) return PD_ERR_ARG;
    ENTER(b);
    const int32_t* dm = nullptr;
    if (mask) {
        std::vector<int32_t> m(b->n); for (int i = 0; i < b->n; ++i) m[i] = mask[i] ? 1 : 0;
        CK(cudaMemcpyAsync(b->dMask, m.data(), (size_t)b->n * 4, cudaMemcpyHostToDevice, b->stream)); CK(cudaStreamSynchronize(b->stream)); dm = b->dMask;
    }
    k_zero_counters<<<grid(b->n, 256), 256, 0, b->stream>>>(b->n, dm, b->dEnvReturn, b->dEnvLen); b->launches++;
    CK(cudaGetLastError()); return PD_OK;
}
int pd_params_bytes(void) { return (int)sizeof(PdCarParams); }
int pd_set_collision_response(pd_batch* b, int on) {
    if (!b) return PD_ERR_ARG;
    ENTER(b);
    b->response = on != 0;
    CK(cudaMemsetAsync(b->dContacts, 0, (size_t)b->n * PD_CONTACT_WORDS * 4, b->stream));
    return PD_OK;
}
int pd_get_contacts(pd_batch* b, int env, float* out, int max_contacts) {
    if (!b || !out || env < 0 || env >= b->n || max_contacts < 0) return -1;
    ENTER(b);
    float rec[PD_CONTACT_WORDS];
    if (cudaMemcpyAsync(rec, b->dContacts + (size_t)env * PD_CONTACT_WORDS, sizeof(rec), cudaMemcpyDeviceToHost, b->stream) != cudaSuccess || cudaStreamSynchronize(b->stream) != cudaSuccess) return -1;
    int n; memcpy(&n, rec, 4);
    if (n < 0) n = 0; if (n > PD_MAX_CONTACTS) n = PD_MAX_CONTACTS;
    for (int i = 0; i < n && i < max_contacts; ++i) memcpy(out + i * 8, rec + 1 + i * 8, 32);
    return n;
}
int pd_set_stream(pd_batch* b, void* stream) {
    if (!b) return PD_ERR_ARG;
    ENTER(b);
    CK(cudaStreamSynchronize(b->stream));
    b->stream = stream ? (cudaStream_t)stream : b->ownStream;
    return PD_OK;
}
int pd_set_scoring_var(pd_batch* b, const char* name, float value) {
    if (!b || !name) return PD_ERR_ARG;
    try { b->car.setScoringVar(name, value); } catch (const std::exception& ex) { b->err = ex.what(); return PD_ERR_ARG; }
    b->paramsDirty = true; return PD_OK;
}
float pd_get_scoring_var(const pd_batch* b, const char* name) { return (b && name) ? b->car.getScoringVar(name) : 0.0f; }

int pd_set_controls(pd_batch* b, const float* controls, const int8_t* gears, int smooth, int on_device) {
    if (!b || !controls) return PD_ERR_ARG;
    const float* c = controls; const int8_t* g = gears;
    if (!on_device) {
        CK(cudaMemcpyAsync(b->dCtl, controls, (size_t)b->n * 5 * 4, cudaMemcpyHostToDevice, b->stream)); c = b->dCtl;
        if (gears) { CK(cudaMemcpyAsync(b->dGears, gears, (size_t)b->n * 3, cudaMemcpyHostToDevice, b->stream)); g = b->dGears; }
    }
    k_set_controls<<<grid(b->n, 256), 256, 0, b->stream>>>(b->dState, b->layout, b->n, c, g, smooth); b->launches++;
    CK(cudaGetLastError()); return PD_OK;
}
int pd_set_actions(pd_batch* b, const float* actions, int on_device) {
    if (!b || !actions) return PD_ERR_ARG;
    const float* a = actions;
    if (!on_device) { CK(cudaMemcpyAsync(b->dAct, actions, (size_t)b->n * 2 * 4, cudaMemcpyHostToDevice, b->stream)); a = b->dAct; }
    k_set_actions<<<grid(b->n, 256), 256, 0, b->stream>>>(b->dState, b->layout, b->n, a, b->envCfg); b->launches++;
    CK(cudaGetLastError()); return PD_OK;
}

int pd_step(pd_batch* b, float dt, int n_ticks) {
    if (!b || n_ticks < 0) return PD_ERR_ARG;
    int rc = sync_params(b); if (rc) return rc;
    for (int t = 0; t < n_ticks; ++t) {
        launch_tick(b, dt, nullptr, EnvIO{});
        b->time += (double)dt; b->lastDt = dt;
    }
    CK(cudaGetLastError()); return PD_OK;
}
double pd_get_time(const pd_batch* b) { return b ? b->time : 0.0; }
int pd_set_time(pd_batch* b, double t) { if (!b) return PD_ERR_ARG; b->time = t; return PD_OK; }
int pd_set_seed(pd_batch* b, uint64_t seed, uint64_t off) { if (!b) return PD_ERR_ARG; b->seed = seed; b->idOffset = off; return PD_OK; }

static int teleport_common(pd_batch* b, const uint8_t* mask, int mode, const float* dist_norm) {
    int rc = sync_params(b); if (rc) return rc;
    const int32_t* dm = nullptr;
    if (mask) {
        std::vector<int32_t> m(b->n); for (int i = 0; i < b->n; ++i) m[i] = mask[i] ? 1 : 0;
        CK(cudaMemcpyAsync(b->dMask, m.data(), (size_t)b->n * 4, cudaMemcpyHostToDevice, b->stream)); CK(cudaStreamSynchronize(b->stream)); dm = b->dMask;
    }
    const float* dd = nullptr;
    if (dist_norm) { CK(cudaMemcpyAsync(b->dDist, dist_norm, (size_t)b->n * 4, cudaMemcpyHostToDevice, b->stream)); dd = b->dDist; }
    k_teleport<<<grid(b->n, PD_BLOCK), PD_BLOCK, 0, b->stream>>>(b->dP, b->dev, b->dState, b->layout, b->n, dm, mode, dd, b->seed, b->idOffset, b->dEpisodeCtr, b->time, 0, b->dPending, b->envCfg); b->launches++;
    CK(cudaGetLastError()); return PD_OK;
}
int pd_teleport_spline(pd_batch* b, const uint8_t* mask, const float* dist_norm) {
    if (!b) return PD_ERR_ARG;
    if (!dist_norm) return teleport_common(b, mask, PD_TELEPORT_START, nullptr);
    return teleport_common(b, mask, PD_TELEPORT_START, dist_norm);
}
int pd_teleport_mode(pd_batch* b, const uint8_t* mask, int mode) {
    if (!b || mode < 0 || mode > 2) return PD_ERR_ARG;
    b->resetMode = mode;
    return teleport_common(b, mask, mode, nullptr);
}

int pd_observe(pd_batch* b) {
    if (!b) return PD_ERR_ARG;
    k_observe<<<grid(b->n, 128), 128, 0, b->stream>>>(b->dState, b->layout, b->n, b->dObs); b->launches++;
    CK(cudaGetLastError()); return PD_OK;
}
const float* pd_obs_device_ptr(pd_batch* b) { return b ? b->dObs : nullptr; }
int pd_get_obs(pd_batch* b, float* out, int to_device) {
    if (!b || !out) return PD_ERR_ARG;
    int rc = pd_observe(b); if (rc) return rc;
    CK(cudaMemcpyAsync(out, b->dObs, (size_t)b->n * PD_OBS_DIM * 4, to_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, b->stream));
    if (!to_device) CK(cudaStreamSynchronize(b->stream));
    return PD_OK;
}
static void dl_deleter(PdDLManagedTensor* t) { delete t; }
void* pd_obs_dlpack(pd_batch* b) {
    if (!b) return nullptr;
    PdDLManagedTensor* t = new PdDLManagedTensor();
    b->dlShape[0] = b->n; b->dlShape[1] = PD_OBS_DIM;
    t->dl_tensor.data = b->dObs; t->dl_tensor.device.device_type = 2 /* kDLCUDA */; t->dl_tensor.device.device_id = b->device;
    t->dl_tensor.ndim = 2; t->dl_tensor.dtype.code = 2 /* kDLFloat */; t->dl_tensor.dtype.bits = 32; t->dl_tensor.dtype.lanes = 1;
    t->dl_tensor.shape = b->dlShape; t->dl_tensor.strides = nullptr; t->dl_tensor.byte_offset = 0;
    t->manager_ctx = b; t->deleter = dl_deleter;
    return t;
}
int pd_get_rewards(pd_batch* b, float* step_reward, float* total_reward, int32_t* flags) {
    if (!b) return PD_ERR_ARG;
    k_rewards<<<grid(b->n, 256), 256, 0, b->stream>>>(b->dState, b->layout, b->n, b->dReward, b->dTotal, b->dFlags); b->launches++;
    CK(cudaGetLastError());
    if (step_reward) CK(cudaMemcpyAsync(step_reward, b->dReward, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (total_reward) CK(cudaMemcpyAsync(total_reward, b->dTotal, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (flags) CK(cudaMemcpyAsync(flags, b->dFlags, (size_t)b->n * 4, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream)); return PD_OK;
}

int pd_env_step(pd_batch* b, const float* actions_dev, float dt, float* obs_dev, float* reward_dev, int32_t* done_dev) {
    if (!b || !actions_dev) return PD_ERR_ARG;
    int rc = sync_params(b); if (rc) return rc;
    const int n = b->n;
    float* rew = reward_dev ? reward_dev : b->dReward; int32_t* done = done_dev ? done_dev : b->dDone;
    /* 1: actions -> controls, one tick, observation, reward / done / statistics -- one launch */
    EnvIO io{}; io.act = actions_dev; io.obs = obs_dev ? obs_dev : b->dObs;
    io.reward = rew; io.done = done; io.envReturn = b->dEnvReturn; io.envLen = b->dEnvLen; io.stats = b->dStats; io.timeAfter = b->time;
    if (b->autoreset == PD_AUTORESET_NEXT_STEP) {
        /* ONE launch per env step: envs that finished at the previous step reset inside this tick kernel */
        io.pending = b->dPending; io.episodeCtr = b->dEpisodeCtr; io.seed = b->seed; io.idOffset = b->idOffset; io.teleportMode = b->resetMode;
        launch_tick(b, dt, nullptr, io);
        b->time += (double)dt; b->lastDt = dt;
        CK(cudaGetLastError()); return PD_OK;
    }
    launch_tick(b, dt, nullptr, io);
    b->time += (double)dt; b->lastDt = dt;
    /* 2 + 3: auto-reset of finished envs: teleport (env.teleport_mode, projectd_env.py:39) with the reset's zero action,
       then one tick of those envs only, refreshing their observation */
    k_teleport<<<grid(n, PD_BLOCK), PD_BLOCK, 0, b->stream>>>(b->dP, b->dev, b->dState, b->layout, n, done, b->resetMode, nullptr, b->seed, b->idOffset, b->dEpisodeCtr, b->time, 1, nullptr, b->envCfg); b->launches++;
    EnvIO io2{}; io2.obs = io.obs;
    launch_tick(b, dt, done, io2);
    CK(cudaGetLastError()); return PD_OK;
}
/* device alias of a page-locked, mapped host buffer (what cudaHostAlloc / torch's pin_memory() hand out on a UVA system), or
 * null when `p` is pageable memory */
static void* mapped_alias(const void* p) {
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
}
int pd_env_step_host(pd_batch* b, const float* actions_host, float dt, float* obs_host, float* reward_host, int32_t* done_host) {
    if (!b || !actions_host) return PD_ERR_ARG;
    const size_t n = (size_t)b->n;
    /* zero-copy path: with the one-launch step (next-step auto-reset) and page-locked host buffers the tick kernel reads the
       actions from, and writes observation / reward / done to, the caller's memory directly over PCIe -- no staging copies, one
       launch (two on odd physics frames) and one stream synchronisation per step.  PD_E2E_ZEROCOPY=0 disables it. */
    if (b->autoreset == PD_AUTORESET_NEXT_STEP && b->zeroCopy && obs_host && reward_host && done_host) {
        if (b->zcKey[0] != actions_host || b->zcKey[1] != obs_host || b->zcKey[2] != reward_host || b->zcKey[3] != done_host) {
            b->zcKey[0] = actions_host; b->zcKey[1] = obs_host; b->zcKey[2] = reward_host; b->zcKey[3] = done_host;
            for (int i = 0; i < 4; ++i) b->zcDev[i] = mapped_alias(b->zcKey[i]);
        }
        if (b->zcDev[0] && b->zcDev[1] && b->zcDev[2] && b->zcDev[3]) {
            int rc = pd_env_step(b, (const float*)b->zcDev[0], dt, (float*)b->zcDev[1], (float*)b->zcDev[2], (int32_t*)b->zcDev[3]); if (rc) return rc;
            CK(cudaStreamSynchronize(b->stream)); return PD_OK;
        }
    }
    CK(cudaMemcpyAsync(b->dAct, actions_host, n * 2 * 4, cudaMemcpyHostToDevice, b->stream));
    int rc = pd_env_step(b, b->dAct, dt, b->dObs, b->dReward, b->dDone); if (rc) return rc;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, b->dObs, n * PD_OBS_DIM * 4, cudaMemcpyDeviceToHost, b->stream));
    if (reward_host) CK(cudaMemcpyAsync(reward_host, b->dReward, n * 4, cudaMemcpyDeviceToHost, b->stream));
    if (done_host) CK(cudaMemcpyAsync(done_host, b->dDone, n * 4, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream)); return PD_OK;
}
int pd_set_autoreset(pd_batch* b, int mode) {
    if (!b || (mode != PD_AUTORESET_SAME_STEP && mode != PD_AUTORESET_NEXT_STEP)) return PD_ERR_ARG;
    b->autoreset = mode;
    CK(cudaMemsetAsync(b->dPending, 0, (size_t)b->n * 4, b->stream));
    return PD_OK;
}
/* profiling aid: per-warp SM cycles of the last full tick launch (only when the batch was created with PD_DEBUG_CLOCKS=1) */
int pd_debug_read_clocks(pd_batch* b, long long* out, int cap) {
    if (!b || !out || !b->dClk) return 0;
    const int blocks = b->layout == PD_LAYOUT_RECORDS ? grid(b->n, 2 * b->quadCpw) : grid(b->n, PD_BLOCK);
    int cnt = blocks * 2;
#if defined(PD_PHASE_CLOCKS)
    cnt = b->nClk; (void)blocks;
#endif
    if (cnt > cap) cnt = cap; if (cnt > b->nClk) cnt = b->nClk;
    if (cudaMemcpyAsync(out, b->dClk, (size_t)cnt * 8, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess || cudaStreamSynchronize(b->stream) != cudaSuccess) return 0;
    return cnt;
}
int pd_env_stats(pd_batch* b, double* out8, int reset) {
    if (!b || !out8) return PD_ERR_ARG;
    CK(cudaMemcpyAsync(out8, b->dStats, 64, cudaMemcpyDeviceToHost, b->stream)); CK(cudaStreamSynchronize(b->stream));
    if (reset) CK(cudaMemsetAsync(b->dStats, 0, 64, b->stream));
    return PD_OK;
}

int pd_get_state(pd_batch* b, int env, uint32_t* record) {
    if (!b || !record || env < 0 || env >= b->n) return PD_ERR_ARG;
    if (b->layout == PD_LAYOUT_RECORDS) CK(cudaMemcpyAsync(record, b->dState + (size_t)env * PD_STATE_STRIDE, PD_STATE_WORDS * 4, cudaMemcpyDeviceToHost, b->stream));
    else CK(cudaMemcpy2DAsync(record, 4, b->dState + state_index_tiled(0, (size_t)env), (size_t)PD_TILE * 4, 4, PD_STATE_WORDS, cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream)); return PD_OK;
}
int pd_set_state(pd_batch* b, int env, const uint32_t* record) {
    if (!b || !record || env < 0 || env >= b->n) return PD_ERR_ARG;
    b->frameKnown = -1;
    CK(cudaMemsetAsync(b->dContacts + (size_t)env * PD_CONTACT_WORDS, 0, PD_CONTACT_WORDS * 4, b->stream));   /* contact joints are not part of the record */
    if (b->layout == PD_LAYOUT_RECORDS) CK(cudaMemcpyAsync(b->dState + (size_t)env * PD_STATE_STRIDE, record, PD_STATE_WORDS * 4, cudaMemcpyHostToDevice, b->stream));
    else CK(cudaMemcpy2DAsync(b->dState + state_index_tiled(0, (size_t)env), (size_t)PD_TILE * 4, record, 4, 4, PD_STATE_WORDS, cudaMemcpyHostToDevice, b->stream));
    CK(cudaStreamSynchronize(b->stream)); return PD_OK;
}
static int pack_common(pd_batch* b, uint32_t* host_buf, int toDevice) {
    const size_t words = (size_t)b->n * PD_STATE_WORDS;
    uint32_t* tmp = nullptr;
    cudaError_t e = cudaMalloc(&tmp, words * 4);
    if (e != cudaSuccess) { b->err = std::string("cudaMalloc: ") + cudaGetErrorString(e); return PD_ERR_CUDA; }
    int rc = PD_OK;
    do {
        if (toDevice && cudaMemcpyAsync(tmp, host_buf, words * 4, cudaMemcpyHostToDevice, b->stream) != cudaSuccess) { rc = PD_ERR_CUDA; break; }
        k_pack<<<(unsigned)((words + 255) / 256), 256, 0, b->stream>>>(b->dState, b->layout, b->n, tmp, toDevice); b->launches++;
        if (!toDevice && cudaMemcpyAsync(host_buf, tmp, words * 4, cudaMemcpyDeviceToHost, b->stream) != cudaSuccess) { rc = PD_ERR_CUDA; break; }
        if (cudaStreamSynchronize(b->stream) != cudaSuccess) rc = PD_ERR_CUDA;
    } while (0);
    if (rc) b->err = std::string("snapshot/restore: ") + cudaGetErrorString(cudaGetLastError());
    cudaFree(tmp);
    return rc;
}
int pd_snapshot(pd_batch* b, uint32_t* host_buf) { if (!b || !host_buf) return PD_ERR_ARG; return pack_common(b, host_buf, 0); }
int pd_restore(pd_batch* b, const uint32_t* host_buf) {
    if (!b || !host_buf) return PD_ERR_ARG;
    b->frameKnown = -1;
    CK(cudaMemsetAsync(b->dContacts, 0, (size_t)b->n * PD_CONTACT_WORDS * 4, b->stream));      /* contact joints are not part of the record: a restored state has none alive */
    return pack_common(b, const_cast<uint32_t*>(host_buf), 1);
}
int pd_get_params(const pd_batch* b, PdCarParams* out) { if (!b || !out) return PD_ERR_ARG; *out = b->car.P; return PD_OK; }
int pd_get_track_info(const pd_batch* b, PdTrackInfo* out) { if (!b || !out) return PD_ERR_ARG; *out = b->track.info; return PD_OK; }

int pd_get_car_state(pd_batch* b, int env, void* outv) {
    if (!b || !outv) return PD_ERR_ARG;
    std::vector<uint32_t> rec(PD_STATE_WORDS);
    int rc = pd_get_state(b, env, rec.data()); if (rc) return rc;
    SVFlat sv = sv_flat(rec.data());
    PdCarStateOut s; memset(&s, 0, sizeof(s));
    CarS c; load_car(sv, c);
    Body C; load_body(sv, PD_BODY_CHASSIS, C);
    s.carId = 0; s.simId = env; s.timestamp = (float)(b->time - b->lastDt);
    s.steer = c.ctlSteer; s.clutch = c.ctlClutch; s.brake = c.ctlBrake; s.handBrake = c.ctlHandBrake; s.gas = c.ctlGas;
    s.isShifterSupported = (int8_t)b->car.P.drivetrain.isShifterSupported; s.requestedGearIndex = (int8_t)c.ctlRequestedGear; s.gearUp = (int8_t)c.ctlGearUp; s.gearDn = (int8_t)c.ctlGearDn;
    s.collisionFlag = c.collisionFlag; s.outOfTrackFlag = c.outOfTrackFlag; s.trackPointId = c.nearestTrackPointId;
    s.lastTrackPointTimestamp = c.lastTrackPointTimestamp; s.trackLocation = c.trackLocation; s.bodyVsTrack = c.bodyVsTrack; s.velocityVsTrack = c.velocityVsTrack;
    s.engineRPM = car_engine_rpm(c); s.speedMS = c.speed; s.gear = c.currentGear; s.gearGrinding = c.isGearGrinding ? 1 : 0;
    auto put_matrix = [](float* m, const Frame& f) {
        m[0] = f.ax.x; m[1] = f.ax.y; m[2] = f.ax.z; m[3] = 0; m[4] = f.ay.x; m[5] = f.ay.y; m[6] = f.ay.z; m[7] = 0;
        m[8] = f.az.x; m[9] = f.az.y; m[10] = f.az.z; m[11] = 0; m[12] = f.p.x; m[13] = f.p.y; m[14] = f.p.z; m[15] = 1.0f; };
    put_matrix(s.bodyMatrix, C.fr);
    s.bodyPos[0] = C.fr.p.x; s.bodyPos[1] = C.fr.p.y; s.bodyPos[2] = C.fr.p.z;
    { /* mat44f::getEulerAngles (Core/Math.cpp:60-86) */
        const float* M = s.bodyMatrix; const float M11 = M[0], M12 = M[1], M21 = M[4], M22 = M[5], M31 = M[8], M32 = M[9], M33 = M[10];
        float rx = atan2f(-M31, M33); float v7 = 1, v8 = M32;
        if (v8 > 1.0 || (v7 = -1, v8 < -1.0)) v8 = v7;
        const float ry = asinf(v8); float v10, v11;
        if (M12 == 0.0f && M22 == 0.0f) { v11 = M21; v10 = M11; rx = 0.0f; } else { v11 = -M12; v10 = M22; }
        const float rz = atan2f(v11, v10);
        s.bodyEuler[0] = ry * -57.295779513082323f; s.bodyEuler[1] = rx * -57.295779513082323f; s.bodyEuler[2] = rz * -57.295779513082323f;
    }
    s.accG[0] = c.accGX; s.accG[1] = c.accGY; s.accG[2] = c.accGZ;
    s.velocity[0] = C.v.x; s.velocity[1] = C.v.y; s.velocity[2] = C.v.z;
    { const V3 lv = irot(C.fr, C.v), lw = irot(C.fr, C.w); s.localVelocity[0] = lv.x; s.localVelocity[1] = lv.y; s.localVelocity[2] = lv.z; s.localAngularVelocity[0] = lw.x; s.localAngularVelocity[1] = lw.y; s.localAngularVelocity[2] = lw.z; }
    s.angularVelocity[0] = C.w.x; s.angularVelocity[1] = C.w.y; s.angularVelocity[2] = C.w.z;
    for (int w = 0; w < 4; ++w) {
        Frame hf;
        const int topo = b->car.P.topology;
        if (w < 2) { Body H; load_body(sv, PD_BODY_HUB0 + 2 * w, H); hf = PD_TOPO_FRONT_DW(topo) ? dw_hub_frame(b->car.P.dw[w], H) : strut_hub_frame(b->car.P.strut[w], H); }
        else if (PD_TOPO_REAR_DW(topo)) { Body H; load_body(sv, PD_BODY_HUB2 + (w - 2), H); hf = dw_hub_frame(b->car.P.dw[w], H); }
        else { Body A; load_body(sv, PD_BODY_AXLE, A); hf = axle_hub_frame(b->car.P.axle, A, w - 2); }
        put_matrix(s.hubMatrix[w], hf);
        const int o = PD_OFF_TYRE(w);
        s.tyreContacts[w][0] = sv.f(o + PD_TYRE_o_contactX); s.tyreContacts[w][1] = sv.f(o + PD_TYRE_o_contactY); s.tyreContacts[w][2] = sv.f(o + PD_TYRE_o_contactZ);
        s.tyreLoad[w] = sv.f(o + PD_TYRE_o_load); s.tyreAngularSpeed[w] = sv.f(o + PD_TYRE_o_angularVelocity);
        s.tyreSlipRatio[w] = sv.f(o + PD_TYRE_o_slipRatio); s.tyreNdSlip[w] = sv.f(o + PD_TYRE_o_ndSlip);
    }
    for (int i = 0; i < b->car.P.nProbes && i < 10; ++i) s.probes[i] = c.probes[i];
    for (int i = 0; i < 5; ++i) s.lookAhead[i] = c.lookAhead[i];
    s.stepReward = c.stepReward; s.totalReward = c.totalReward;
    memcpy(outv, &s, sizeof(s)); return PD_OK;
}

int pd_raycast(pd_batch* b, int n, const float* rays, float* out) {
    if (!b || n < 0 || !rays || !out) return PD_ERR_ARG;
    if (n == 0) return PD_OK;
    float *dr = nullptr, *dout = nullptr;
    CK(cudaMalloc(&dr, (size_t)n * 7 * 4)); CK(cudaMalloc(&dout, (size_t)n * 8 * 4));
    cudaMemcpyAsync(dr, rays, (size_t)n * 7 * 4, cudaMemcpyHostToDevice, b->stream);
    k_raycast<<<grid(n, 128), 128, 0, b->stream>>>(b->dev, n, dr, dout); b->launches++;
    cudaMemcpyAsync(out, dout, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost, b->stream);
    cudaError_t e = cudaStreamSynchronize(b->stream);
    cudaFree(dr); cudaFree(dout);
    if (e != cudaSuccess) { b->err = cudaGetErrorString(e); return PD_ERR_CUDA; }
    return PD_OK;
}

int pd_sync(pd_batch* b) { if (!b) return PD_ERR_ARG; CK(cudaStreamSynchronize(b->stream)); CK(cudaGetLastError()); return PD_OK; }
void* pd_stream(pd_batch* b) { return b ? (void*)b->stream : nullptr; }
const char* pd_tick_kernel(const pd_batch* b) { return !b ? "" : (b->layout == PD_LAYOUT_RECORDS ? "k_tick_quad" : "k_tick"); }
int pd_topology(const pd_batch* b) { return b ? b->car.P.topology : -1; }
int pd_bvh_info(const pd_batch* b, int* built_on_device, int* n_nodes, int* depth) {
    if (!b) return PD_ERR_ARG;
    if (built_on_device) *built_on_device = b->bvhOnDevice;
    if (n_nodes) *n_nodes = b->dev.info.nNodes;
    if (depth) *depth = b->bvhDepth;
    return PD_OK;
}
const char* pd_tick_kernel_instance(const pd_batch* b) {
    if (!b) return "";
    if (b->layout != PD_LAYOUT_RECORDS) return b->car.P.topology == PD_TOPO_STRUT_DW ? "k_tick<strut,dwb>" : (b->car.P.topology == PD_TOPO_DW_DW ? "k_tick<dwb,dwb>" : (b->serialWide ? "k_tick/255" : "k_tick"));
    if (b->car.P.topology == PD_TOPO_STRUT_DW) return b->quadCpw == 4 ? "k_tick_quad<4,strut,dwb>" : "k_tick_quad<8,strut,dwb>";
    if (b->car.P.topology == PD_TOPO_DW_DW) return b->quadCpw == 4 ? "k_tick_quad<4,dwb,dwb>" : "k_tick_quad<8,dwb,dwb>";
    return b->quadCpw == 2 ? "k_tick_quad<2>" : b->quadCpw == 4 ? "k_tick_quad<4>" : "k_tick_quad<8>";
}
uint64_t pd_launch_count(const pd_batch* b) { return b ? b->launches : 0; }

} /* extern "C" */
