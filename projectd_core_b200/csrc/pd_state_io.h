/*
 * pd_state_io.h -- access to the per-car state record (include/pd_state.h) in its device layouts.
 */
#pragma once
#include "pd_math.h"
#include "../../include/pd_state.h"
#include "../../include/pd_params.h"
#include <string.h>
#include <stddef.h>

namespace pd {

PD_HD float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
PD_HD uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
PD_HD double u2d(uint32_t lo, uint32_t hi) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    uint64_t v = ((uint64_t)hi << 32) | lo; double d; memcpy(&d, &v, 8); return d;
#endif
}
PD_HD void d2u(double d, uint32_t& lo, uint32_t& hi) {
#if defined(__CUDA_ARCH__)
    lo = (uint32_t)__double2loint(d); hi = (uint32_t)__double2hiint(d);
#else
    uint64_t v; memcpy(&v, &d, 8); lo = (uint32_t)v; hi = (uint32_t)(v >> 32);
#endif
}

/* Views of one env inside the state buffer.  Two device layouts exist (chosen per batch at creation):
 *
 *  RECORDS ("array of records", 4-lanes-per-car kernel): env e is the contiguous record
 *      state[e * PD_STATE_STRIDE .. + PD_STATE_WORDS).  A block moves its 16 records between HBM and shared
 *      memory with ONE bulk async copy each way and works on the shared-memory copy (view stride 1).
 *  TILED  ("tiled structure of arrays", thread-per-car kernel): envs are grouped in tiles of 32; word w of
 *      env e lives at state[((e / 32) * PD_STATE_WORDS + w) * 32 + (e % 32)].  A warp's accesses to one word
 *      of 32 consecutive envs are one 128-byte line and, the word index being a compile-time constant almost
 *      everywhere, every access is a load/store with an IMMEDIATE offset from one per-thread base pointer.
 *
 *  SVT<STRIDE>: view with a compile-time word stride (SVT<1> = one flat record, SVT<PD_TILE> = tiled);
 *  SVR: view with a run-time stride, for the small kernels that serve both layouts. */
#define PD_TILE 32
enum { PD_LAYOUT_TILED = 0, PD_LAYOUT_RECORDS = 1 };
#if defined(__CUDA_ARCH__)
/* address-space hints: a tiled view is only ever built over the global state buffer, a stride-1 view on the device
 * over the shared-memory staging copy (no hint there: nvcc 12.9 miscompiles __builtin_assume(__isShared()) on a pointer
 * into dynamic shared memory -- the kernel body is dropped); without the hint the accesses inside non-inlined
 * functions are generic LD/ST */
#define PD_ASSUME_SPACE(p) do { if (STRIDE != 1) __builtin_assume(__isGlobal(p)); } while (0)
#else
#define PD_ASSUME_SPACE(p) do { } while (0)
#endif
template <int STRIDE> struct SVT {
    uint32_t* s;        /* &state[word 0 of this env] */
    static constexpr int stride = STRIDE;
    PD_HD explicit SVT(uint32_t* base) : s(base) {}
    PD_HD float f(int w) const { PD_ASSUME_SPACE(s); return u2f(s[w * STRIDE]); }
    PD_HD int i(int w) const { PD_ASSUME_SPACE(s); return (int)s[w * STRIDE]; }
    PD_HD double d(int w) const { PD_ASSUME_SPACE(s); return u2d(s[w * STRIDE], s[(w + 1) * STRIDE]); }
    PD_HD void f(int w, float v) const { PD_ASSUME_SPACE(s); s[w * STRIDE] = f2u(v); }
    PD_HD void i(int w, int v) const { PD_ASSUME_SPACE(s); s[w * STRIDE] = (uint32_t)v; }
    PD_HD void d(int w, double v) const { PD_ASSUME_SPACE(s); uint32_t lo, hi; d2u(v, lo, hi); s[w * STRIDE] = lo; s[(w + 1) * STRIDE] = hi; }
};
struct SVR {
    uint32_t* s; int stride;
    PD_HD SVR(uint32_t* base, int st) : s(base), stride(st) {}
    PD_HD float f(int w) const { return u2f(s[w * stride]); }
    PD_HD int i(int w) const { return (int)s[w * stride]; }
    PD_HD double d(int w) const { return u2d(s[w * stride], s[(w + 1) * stride]); }
    PD_HD void f(int w, float v) const { s[w * stride] = f2u(v); }
    PD_HD void i(int w, int v) const { s[w * stride] = (uint32_t)v; }
    PD_HD void d(int w, double v) const { uint32_t lo, hi; d2u(v, lo, hi); s[w * stride] = lo; s[(w + 1) * stride] = hi; }
};
typedef SVT<PD_TILE> SVTile;
typedef SVT<1> SVFlat;
PD_HD size_t state_index_tiled(int w, size_t e) { return ((e / PD_TILE) * PD_STATE_WORDS + (size_t)w) * PD_TILE + (e % PD_TILE); }
PD_HD size_t state_index(int layout, int w, size_t e) { return layout == PD_LAYOUT_RECORDS ? e * PD_STATE_STRIDE + (size_t)w : state_index_tiled(w, e); }
PD_HD size_t state_alloc_words(int layout, size_t n) { return layout == PD_LAYOUT_RECORDS ? n * PD_STATE_STRIDE : ((n + PD_TILE - 1) / PD_TILE * PD_TILE) * PD_STATE_WORDS; }
PD_HD SVTile sv_tiled(uint32_t* state, size_t e) { return SVTile(state + state_index_tiled(0, e)); }
PD_HD SVFlat sv_flat(uint32_t* rec) { return SVFlat(rec); }
PD_HD SVR sv_env(int layout, uint32_t* state, size_t e) { return layout == PD_LAYOUT_RECORDS ? SVR(state + e * PD_STATE_STRIDE, 1) : SVR(state + state_index_tiled(0, e), PD_TILE); }

/* ---- typed mirrors of the X-macro lists ---- */
#define PD__DECL_F(name) float name;
#define PD__DECL_I(name) int name;
#define PD__DECL_D(name) double name;
#define PD__DECL(kind, name) PD__DECL_##kind(name)

struct TyreS {
    PD_TYRE_FIELDS(PD__DECL)
    float T[PD_THERMAL_PATCHES];
};
struct CarS {
    PD_CAR_FIELDS(PD__DECL)
    float probes[PD_MAX_PROBES];
    float lookAhead[PD_LOOKAHEAD];
};

/* A stride-1 view IS a record in memory with the layout of the typed mirrors (doubles first in each group, groups on
 * even word offsets, see include/pd_state.h), so its tyre / car parts can be used in place instead of being copied. */
#define PD__CHK_T(kind, name) static_assert(offsetof(TyreS, name) == 4 * PD_TYRE_o_##name, "TyreS layout != record layout: " #name);
#define PD__CHK_C(kind, name) static_assert(offsetof(CarS, name) == 4 * PD_CAR_o_##name, "CarS layout != record layout: " #name);
PD_TYRE_FIELDS(PD__CHK_T)
PD_CAR_FIELDS(PD__CHK_C)
static_assert(offsetof(TyreS, T) == 4 * PD_TYRE_SCALAR_WORDS && sizeof(TyreS) == 4 * PD_TYRE_WORDS, "TyreS size");
static_assert(offsetof(CarS, probes) == 4 * PD_CAR_SCALAR_WORDS && offsetof(CarS, lookAhead) == 4 * (PD_CAR_SCALAR_WORDS + PD_MAX_PROBES), "CarS arrays");
static_assert(PD_OFF_TYRE(0) % 2 == 0 && PD_TYRE_WORDS % 2 == 0 && PD_OFF_CAR % 2 == 0 && PD_STATE_STRIDE % 2 == 0, "8-byte alignment of the double fields");
#ifndef PD_GRID_IN_PLACE
#define PD_GRID_IN_PLACE 1      /* tiled view: the tyres' thermal grids are swept in the state buffer, not in a local copy (pd_car.h GridSV) */
#endif
template <class SVX> struct sv_traits { static constexpr bool in_place = false; static constexpr bool grid_in_place = false; };
template <> struct sv_traits<SVT<1> > { static constexpr bool in_place = true; static constexpr bool grid_in_place = false; };
template <> struct sv_traits<SVT<PD_TILE> > { static constexpr bool in_place = false; static constexpr bool grid_in_place = PD_GRID_IN_PLACE != 0; };
PD_HD TyreS* tyre_in_place(const SVT<1>& sv, int w) { return reinterpret_cast<TyreS*>(sv.s + PD_OFF_TYRE(w)); }
PD_HD CarS* car_in_place(const SVT<1>& sv) { return reinterpret_cast<CarS*>(sv.s + PD_OFF_CAR); }
template <class SVX> PD_HD TyreS* tyre_in_place(const SVX&, int) { return nullptr; }
template <class SVX> PD_HD CarS* car_in_place(const SVX&) { return nullptr; }

#define PD__LD_F(pre, name) t.name = sv.f(o + pre##name);
#define PD__LD_I(pre, name) t.name = sv.i(o + pre##name);
#define PD__LD_D(pre, name) t.name = sv.d(o + pre##name);
#define PD__ST_F(pre, name) sv.f(o + pre##name, t.name);
#define PD__ST_I(pre, name) sv.i(o + pre##name, t.name);
#define PD__ST_D(pre, name) sv.d(o + pre##name, t.name);
#define PD__LDT(kind, name) PD__LD_##kind(PD_TYRE_o_, name)
#define PD__STT(kind, name) PD__ST_##kind(PD_TYRE_o_, name)
#define PD__LDC(kind, name) PD__LD_##kind(PD_CAR_o_, name)
#define PD__STC(kind, name) PD__ST_##kind(PD_CAR_o_, name)

template <class SVX> PD_HD void load_tyre(const SVX& sv, int w, TyreS& t) {
    const int o = PD_OFF_TYRE(w);
    PD_TYRE_FIELDS(PD__LDT)
    if constexpr (!sv_traits<SVX>::grid_in_place) {
        PD_UNROLL4
        for (int p = 0; p < PD_THERMAL_PATCHES; ++p) t.T[p] = sv.f(PD_OFF_TYRE_PATCH(w) + p);
    }
}
template <class SVX> PD_HD void store_tyre(const SVX& sv, int w, const TyreS& t) {
    const int o = PD_OFF_TYRE(w);
    PD_TYRE_FIELDS(PD__STT)
    if constexpr (!sv_traits<SVX>::grid_in_place) {
        PD_UNROLL4
        for (int p = 0; p < PD_THERMAL_PATCHES; ++p) sv.f(PD_OFF_TYRE_PATCH(w) + p, t.T[p]);
    }
}
template <class SVX> PD_HD void load_car(const SVX& sv, CarS& t) {
    const int o = PD_OFF_CAR;
    PD_CAR_FIELDS(PD__LDC)
    for (int p = 0; p < PD_MAX_PROBES; ++p) t.probes[p] = sv.f(PD_OFF_PROBES + p);
    for (int p = 0; p < PD_LOOKAHEAD; ++p) t.lookAhead[p] = sv.f(PD_OFF_LOOKAHEAD + p);
}
template <class SVX> PD_HD void store_car(const SVX& sv, const CarS& t) {
    const int o = PD_OFF_CAR;
    PD_CAR_FIELDS(PD__STC)
    for (int p = 0; p < PD_MAX_PROBES; ++p) sv.f(PD_OFF_PROBES + p, t.probes[p]);
    for (int p = 0; p < PD_LOOKAHEAD; ++p) sv.f(PD_OFF_LOOKAHEAD + p, t.lookAhead[p]);
}

/* rigid body working copy: dxBody state + force / torque accumulators + mass data */
struct Body {
    Frame fr;      /* pos + axes (R) */
    Quat q;
    V3 v, w;       /* lvel, avel */
    V3 F, T;       /* facc, tacc */
    float mass;
    V3 I;          /* diagonal body-frame inertia */
};
template <class SVX> PD_HD void load_body(const SVX& sv, int b, Body& B) {
    const int o = PD_OFF_BODY(b);
    B.fr.p = v3(sv.f(o + PD_BODY_o_px), sv.f(o + PD_BODY_o_py), sv.f(o + PD_BODY_o_pz));
    B.q.w = sv.f(o + PD_BODY_o_qw); B.q.x = sv.f(o + PD_BODY_o_qx); B.q.y = sv.f(o + PD_BODY_o_qy); B.q.z = sv.f(o + PD_BODY_o_qz);
    B.fr.ax = v3(sv.f(o + PD_BODY_o_axx), sv.f(o + PD_BODY_o_axy), sv.f(o + PD_BODY_o_axz));
    B.fr.ay = v3(sv.f(o + PD_BODY_o_ayx), sv.f(o + PD_BODY_o_ayy), sv.f(o + PD_BODY_o_ayz));
    B.fr.az = v3(sv.f(o + PD_BODY_o_azx), sv.f(o + PD_BODY_o_azy), sv.f(o + PD_BODY_o_azz));
    B.v = v3(sv.f(o + PD_BODY_o_vx), sv.f(o + PD_BODY_o_vy), sv.f(o + PD_BODY_o_vz));
    B.w = v3(sv.f(o + PD_BODY_o_wx), sv.f(o + PD_BODY_o_wy), sv.f(o + PD_BODY_o_wz));
    B.F = v3(0, 0, 0); B.T = v3(0, 0, 0);
}
template <class SVX> PD_HD void store_body(const SVX& sv, int b, const Body& B) {
    const int o = PD_OFF_BODY(b);
    sv.f(o + PD_BODY_o_px, B.fr.p.x); sv.f(o + PD_BODY_o_py, B.fr.p.y); sv.f(o + PD_BODY_o_pz, B.fr.p.z);
    sv.f(o + PD_BODY_o_qw, B.q.w); sv.f(o + PD_BODY_o_qx, B.q.x); sv.f(o + PD_BODY_o_qy, B.q.y); sv.f(o + PD_BODY_o_qz, B.q.z);
    sv.f(o + PD_BODY_o_vx, B.v.x); sv.f(o + PD_BODY_o_vy, B.v.y); sv.f(o + PD_BODY_o_vz, B.v.z);
    sv.f(o + PD_BODY_o_wx, B.w.x); sv.f(o + PD_BODY_o_wy, B.w.y); sv.f(o + PD_BODY_o_wz, B.w.z);
    sv.f(o + PD_BODY_o_axx, B.fr.ax.x); sv.f(o + PD_BODY_o_axy, B.fr.ax.y); sv.f(o + PD_BODY_o_axz, B.fr.ax.z);
    sv.f(o + PD_BODY_o_ayx, B.fr.ay.x); sv.f(o + PD_BODY_o_ayy, B.fr.ay.y); sv.f(o + PD_BODY_o_ayz, B.fr.ay.z);
    sv.f(o + PD_BODY_o_azx, B.fr.az.x); sv.f(o + PD_BODY_o_azy, B.fr.az.y); sv.f(o + PD_BODY_o_azz, B.fr.az.z);
}

/* 0 when position, velocities and q.w of the body are all finite, NaN otherwise: x - x is 0 for a finite x and NaN for an infinity or a NaN, and
 * NaN survives the sum -- the same predicate as ten finitef() tests, without their ten dependent branches */
PD_HD float body_nonfinite_acc(const Body& b) {
    return ((b.fr.p.x - b.fr.p.x) + (b.fr.p.y - b.fr.p.y)) + ((b.fr.p.z - b.fr.p.z) + (b.v.x - b.v.x)) + ((b.v.y - b.v.y) + (b.v.z - b.v.z)) + ((b.w.x - b.w.x) + (b.w.y - b.w.y)) + ((b.w.z - b.w.z) + (b.q.w - b.q.w));
}

/* ---- body force API (RigidBodyODE.cpp:184-270 -> ODE dBody*) ---- */
PD_HD V3 body_point_vel(const Body& b, V3 p) { return b.v + cross(b.w, p - b.fr.p); }            /* dBodyGetPointVel */
PD_HD V3 body_rel_point_vel(const Body& b, V3 prel) { return b.v + cross(b.w, rot(b.fr, prel)); } /* dBodyGetRelPointVel */
PD_HD void add_force_at_pos(Body& b, V3 f, V3 p) { b.F += f; b.T += cross(p - b.fr.p, f); }       /* dBodyAddForceAtPos */
PD_HD void add_force_at_rel_pos(Body& b, V3 f, V3 prel) { b.F += f; b.T += cross(rot(b.fr, prel), f); }
PD_HD void add_rel_force_at_rel_pos(Body& b, V3 frel, V3 prel) { add_force_at_rel_pos(b, rot(b.fr, frel), prel); }
PD_HD void add_torque(Body& b, V3 t) { b.T += t; }
PD_HD void add_rel_torque(Body& b, V3 t) { b.T += rot(b.fr, t); }
PD_HD void body_stop(Body& b) { b.v = v3(0, 0, 0); b.w = v3(0, 0, 0); b.F = v3(0, 0, 0); b.T = v3(0, 0, 0); }

/* Core/Curve.cpp:94-115 */
PD_HD float curve_value(const PdCurve& c, float ref) {
    if (c.n <= 0) return 0.0f;
    if (ref <= c.ref[0]) return c.val[0];
    for (int id = 1; id < c.n; ++id) {
        if (ref <= c.ref[id])
            return (((c.val[id] - c.val[id - 1]) * (ref - c.ref[id - 1])) / (c.ref[id] - c.ref[id - 1])) + c.val[id - 1];
    }
    return c.val[c.n - 1];
}

} // namespace pd
