/*
 * tests/hostsim/hostsim.cpp -- DEBUGGING AID, test-only (never shipped, never loaded by the product).
 *
 * Compiles the product's device functions (projectd_core_b200/csrc/pd_tick.h, PD_HD = plain inline under
 * g++) for the host so that the kernels' arithmetic can be single-stepped against the oracle on a box
 * without a GPU.  The product path is the CUDA library (libpd_b200.so); it has no CPU fallback and does not
 * link this file.
 */
#include "../../projectd_core_b200/csrc/pd_quad.h"
#include <pthread.h>
#include <vector>
#include <thread>
#include "../../projectd_core_b200/csrc/host/pd_host.h"
#include <cstdio>

struct HS {
    pdh::CarModel car;
    pdh::TrackModel track;
    pd::TrackDev dev;
    void bind() {
        dev.nodes = (const pd::BvhNode*)track.nodes.data(); dev.tris = track.tris.data(); dev.triSurf = track.triSurf.data();
        dev.surfaces = track.surfaces.data(); dev.fat = track.fat.data(); dev.splineXYZ = track.splineXYZ.data(); dev.splineDist = track.splineDist.data();
        dev.segStart = track.segStart.data(); dev.segItems = track.segItems.data(); dev.ptStart = track.ptStart.data(); dev.ptItems = track.ptItems.data(); dev.segRec = track.segRec.data(); dev.ptRec = track.ptRec.data(); dev.grid = track.grid; dev.segGrid = track.segGrid;
        dev.colStart = track.colStart.data(); dev.colItems = track.colItems.data(); dev.colGrid = track.colGrid;
        dev.triRaw = track.triRaw.data(); dev.collStart = track.collStart.data(); dev.collItems = track.collItems.data(); dev.collPlane = track.collPlane.data(); dev.collCell = track.collCell.data(); dev.collRec = track.collRec.data(); dev.collGrid = track.collGrid; dev.hullTables = nullptr;
        dev.info = track.info;
    }
};

/* host stand-in for the GPU's quad shuffles: the four lanes of a car run as four threads that meet at a barrier */
struct QuadShared { float slot[4][3]; int islot[4]; pthread_barrier_t bar; };
struct QuadHost {
    int lane; QuadShared* sh;
    int half = 0, nhalf = 1;
    float peer(float v) { return v; }
    int peer(int v) { return v; }
    void sync() { pthread_barrier_wait(&sh->bar); }
    float get(float v, int src) { sh->slot[lane][0] = v; sync(); float r = sh->slot[src][0]; sync(); return r; }
    int get(int v, int src) { sh->islot[lane] = v; sync(); int r = sh->islot[src]; sync(); return r; }
    pd::V3 get(pd::V3 v, int src) { sh->slot[lane][0] = v.x; sh->slot[lane][1] = v.y; sh->slot[lane][2] = v.z; sync(); pd::V3 r = pd::v3(sh->slot[src][0], sh->slot[src][1], sh->slot[src][2]); sync(); return r; }
    float sum(float v) { /* the GPU butterfly: (l0+l1)+(l2+l3) */
        sh->slot[lane][0] = v; sync(); float a = sh->slot[lane][0] + sh->slot[lane ^ 1][0]; sync();
        sh->slot[lane][0] = a; sync(); float r = sh->slot[lane][0] + sh->slot[lane ^ 2][0]; sync(); return r; }
    pd::V3 sum(pd::V3 v) { return pd::v3(sum(v.x), sum(v.y), sum(v.z)); }
    bool all(bool p) { sh->islot[lane] = p ? 1 : 0; sync(); bool r = sh->islot[0] && sh->islot[1] && sh->islot[2] && sh->islot[3]; sync(); return r; }
};

extern "C" {
void* hs_create(const char* base, const char* track, const char* car) {
    HS* h = new HS();
    try { pdh::load_car(base, car, h->car); pdh::load_track(base, track, h->track); h->bind(); }
    catch (const std::exception& e) { fprintf(stderr, "[hostsim] %s\n", e.what()); delete h; return nullptr; }
    return h;
}
void hs_destroy(void* h) { delete (HS*)h; }
void hs_set_tune(void* h, const char* name, float v) { ((HS*)h)->car.setTune(name, v); }
void hs_set_scoring_var(void* h, const char* name, float v) { ((HS*)h)->car.setScoringVar(name, v); }
void hs_set_assists(void* hv, int ac, int as, int ab) { auto& A = ((HS*)hv)->car.P.assists; A.acUseAutoOnStart = ac; A.acUseAutoOnChange = ac; A.asIsActive = as; A.blipIsActive = ab; }
int hs_params_bytes() { return (int)sizeof(PdCarParams); }
int hs_offset_autoshift_rpm() { return (int)offsetof(PdCarParams, assists.asChangeUpRpm); }   /* asChangeUpRpm, asChangeDnRpm: resolved lazily by the reference (AutoShifter.cpp:38-53), at load by the loader */
void hs_get_params(void* h, PdCarParams* out) { *out = ((HS*)h)->car.P; }
void hs_set_params(void* h, const PdCarParams* in) { ((HS*)h)->car.P = *in; }
void hs_get_track_info(void* h, PdTrackInfo* out) { *out = ((HS*)h)->track.info; }
void hs_get_spline_nodes(void* hv, float* xyz, float* dist) { HS* h = (HS*)hv; memcpy(xyz, h->track.splineXYZ.data(), h->track.splineXYZ.size() * 4); memcpy(dist, h->track.splineDist.data(), h->track.splineDist.size() * 4); }
void hs_tick(void* hv, uint32_t* rec, float dt, double time) {
    HS* h = (HS*)hv; pd::SVFlat sv = pd::sv_flat(rec); float scr[PD_GSCR_WORDS];
    switch (h->car.P.topology) {      /* the same compile-time instances the thread-per-car kernel has */
    case PD_TOPO_STRUT_DW: pd::car_tick<1, 1, PD_TOPO_STRUT_DW>(h->car.P, h->dev, sv, dt, time, scr, scr + PD_GSCR_ROWS_WORDS); break;
    case PD_TOPO_DW_DW: pd::car_tick<1, 1, PD_TOPO_DW_DW>(h->car.P, h->dev, sv, dt, time, scr, scr + PD_GSCR_ROWS_WORDS); break;
    default: pd::car_tick<1, 1>(h->car.P, h->dev, sv, dt, time, scr, scr + PD_GSCR_ROWS_WORDS); break;
    }
}
void hs_tick_quad(void* hv, uint32_t* rec, float dt, double time) {
    /* the GPU lanes of a quad share one record and run converged; host threads do not, so every "lane" gets a private
       copy of the record and the parts each lane owns are merged afterwards:
       lane 0: chassis, hub0, strut0, tyre 0, car part;  lane 1: hub1, strut1, tyre 1;  lane 2: axle, tyre 2;  lane 3: tank, tyre 3 */
    HS* h = (HS*)hv; QuadShared sh; pthread_barrier_init(&sh.bar, nullptr, 4);
    std::vector<uint32_t> copy[4];
    for (int l = 0; l < 4; ++l) copy[l].assign(rec, rec + PD_STATE_WORDS);
    std::thread th[4];
    const int topo = h->car.P.topology;
    for (int l = 0; l < 4; ++l) th[l] = std::thread([&, l]() {
        QuadHost ex{l, &sh}; pd::SVFlat sv = pd::sv_flat(copy[l].data()); float scr[PD_GSCR_WORDS];
        if (topo == PD_TOPO_STRUT_DW) pd::car_tick_quad<1, 1, PD_TOPO_STRUT_DW>(h->car.P, h->dev, sv, dt, time, ex, scr, scr + PD_GSCR_ROWS_WORDS);
        else if (topo == PD_TOPO_DW_DW) pd::car_tick_quad<1, 1, PD_TOPO_DW_DW>(h->car.P, h->dev, sv, dt, time, ex, scr, scr + PD_GSCR_ROWS_WORDS);
        else pd::car_tick_quad<1, 1>(h->car.P, h->dev, sv, dt, time, ex, scr, scr + PD_GSCR_ROWS_WORDS);
    });
    for (int l = 0; l < 4; ++l) th[l].join();
    pthread_barrier_destroy(&sh.bar);
    auto take = [&](int lane, int off, int words) { memcpy(rec + off, copy[lane].data() + off, (size_t)words * 4); };
    take(0, PD_OFF_BODY(PD_BODY_CHASSIS), PD_BODY_WORDS); take(0, PD_OFF_BODY(PD_BODY_HUB0), PD_BODY_WORDS); take(1, PD_OFF_BODY(PD_BODY_HUB1), PD_BODY_WORDS);
    if (!PD_TOPO_FRONT_DW(topo)) { take(0, PD_OFF_BODY(PD_BODY_STRUT0), PD_BODY_WORDS); take(1, PD_OFF_BODY(PD_BODY_STRUT1), PD_BODY_WORDS); }
    take(2, PD_OFF_BODY(PD_BODY_AXLE), PD_BODY_WORDS);          /* the rigid axle, or the LR hub of a double-wishbone rear axle */
    if (PD_TOPO_REAR_DW(topo)) take(3, PD_OFF_BODY(PD_BODY_HUB3), PD_BODY_WORDS);
    take(3, PD_OFF_BODY(PD_BODY_TANK), PD_BODY_WORDS);
    for (int l = 0; l < 4; ++l) take(l, PD_OFF_TYRE(l), PD_TYRE_WORDS);
    take(0, PD_OFF_CAR, PD_CAR_WORDS);
}
void hs_teleport_point(void* hv, uint32_t* rec, int pointId, double time) { HS* h = (HS*)hv; pd::SVFlat sv = pd::sv_flat(rec); pd::car_teleport_to_point(h->car.P, h->dev, sv, pointId, time); }
int hs_point_id_at_distance(void* hv, float d) { return pd::point_id_at_distance(((HS*)hv)->dev, d); }
void hs_raycast(void* hv, int n, const float* in, float* out) {
    HS* h = (HS*)hv;
    for (int i = 0; i < n; ++i) {
        const float* p = in + i * 7; float* q = out + i * 8;
        const bool down = (p[3] == 0.0f && p[4] == -1.0f && p[5] == 0.0f);
        pd::RayHit r = down ? pd::ray_cast_down(h->dev, pd::v3(p[0], p[1], p[2]), p[6]) : pd::ray_cast(h->dev, pd::v3(p[0], p[1], p[2]), pd::v3(p[3], p[4], p[5]), p[6]);
        q[0] = (float)r.hit; q[1] = r.pos.x; q[2] = r.pos.y; q[3] = r.pos.z; q[4] = r.normal.x; q[5] = r.normal.y; q[6] = r.normal.z; q[7] = (float)r.surface;
    }
}
/* probes + nearest point: grid-indexed walk vs the reference's exhaustive loop, at arbitrary poses.
 * in[n][4] = x, y, z, yaw ; walk / brute [n][8] = 7 probe distances + nearest point id */
void hs_probe_compare(void* hv, int n, const float* in, float* walk, float* brute) {
    HS* h = (HS*)hv; const pd::TrackDev& T = h->dev; const PdCarParams& P = h->car.P;
    const int nFat = T.info.nFatPoints;
    for (int i = 0; i < n; ++i) {
        const float* p = in + i * 4; const pd::V3 pos = pd::v3(p[0], p[1], p[2]); const float yaw = p[3];
        const float nearR = P.probeLength[0], nearRSq = nearR * nearR;
        for (int r = 0; r < P.nProbes; ++r) {
            const float dx = P.probeDir[r][0] * cosf(yaw) + P.probeDir[r][2] * sinf(yaw), dz = -P.probeDir[r][0] * sinf(yaw) + P.probeDir[r][2] * cosf(yaw);
            const float bx = pos.x + dx * (P.probeLength[r] * 1.1f), bz = pos.z + dz * (P.probeLength[r] * 1.1f);
            float w = FLT_MAX; const bool ok = pd::probe_walk(T, pos.x, pos.z, bx, bz, pos, nearRSq, w);
            float b = FLT_MAX;
            for (int id = 0; id < nFat; ++id) {
                const PdFatPoint& f = T.fat[id];
                if (!(pd::sqlen(pos - pd::v3(f.best[0], f.best[1], f.best[2])) < nearRSq)) continue;
                const PdFatPoint& g = T.fat[id + 1 < nFat ? id + 1 : 0];
                float ix, iz;
                if (pd::line_intersection(pos.x, pos.z, bx, bz, f.left[0], f.left[2], g.left[0], g.left[2], ix, iz)) { const float ex = pos.x - ix, ez = pos.z - iz; b = pd::tminf(b, sqrtf(ex * ex + ez * ez)); }
                if (pd::line_intersection(pos.x, pos.z, bx, bz, f.right[0], f.right[2], g.right[0], g.right[2], ix, iz)) { const float ex = pos.x - ix, ez = pos.z - iz; b = pd::tminf(b, sqrtf(ex * ex + ez * ez)); }
            }
            walk[i * 8 + r] = ok ? w : -1.0f; brute[i * 8 + r] = b;
        }
        int bp = -1; const bool ok = pd::nearest_point_grid(T, pos, pos, nearRSq, bp);
        walk[i * 8 + 7] = ok ? (float)bp : -1.0f;
        float bd = FLT_MAX; int bb = 0;
        for (int id = 0; id < nFat; ++id) { const PdFatPoint& f = T.fat[id]; const pd::V3 loc = pd::v3(f.best[0], f.best[1], f.best[2]); if (!(pd::sqlen(pos - loc) < nearRSq)) continue; const float d = pd::sqlen(loc - pos); if (bd > d) { bd = d; bb = id; } }
        brute[i * 8 + 7] = (float)bb;
    }
}
void hs_sctm_solve(void* hv, int wheel, int n, const float* in, float* out) {
    HS* h = (HS*)hv;
    for (int i = 0; i < n; ++i) {
        const float* p = in + i * 9; pd::TmIn t;
        t.load = p[0]; t.slipAngleRAD = p[1]; t.slipRatio = p[2]; t.camberRAD = p[3]; t.speed = p[4]; t.u = p[5]; t.cpLength = p[6]; t.pressureRatio = p[7]; t.blister = p[8]; t.grain = 0;
        pd::TmOut o = pd::sctm_solve(h->car.P.tyre[wheel], t);
        float* q = out + i * 7; q[0] = o.Fy; q[1] = o.Fx; q[2] = o.Mz; q[3] = o.trail; q[4] = o.ndSlip; q[5] = o.Dy; q[6] = o.Dx;
    }
}
}
