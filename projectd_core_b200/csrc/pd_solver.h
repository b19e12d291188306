/*
 * pd_solver.h -- the per-car rigid-body / joint solve that ODE's dWorldStep performs in the reference
 * (Physics/ODE/PhysicsEngineODE.cpp:216-224, semantics of ODE 0.16.3 step.cpp
 * dxStepIsland restated in oracle/ode_restate/ode_core.h).
 *
 * Demo-car island: 7 bodies, 33 bilateral rows, no inequality rows  ->  one linear solve
 *     (J M^-1 J^T + CFM/h) lambda = c/h - J (v/h + M^-1 f)
 * per car per tick.  The reference factors the dense 33x33 matrix.  Here the island's STAR topology is
 * used instead: every joint group touches only its own bodies and the chassis,
 *     group 0  tank            fixed(tank, chassis)                                   6 rows, own {T}
 *     group 1  LF strut        3 x dball(chassis, hub) + slider(strut, hub) + ball(chassis, strut)   11 rows, own {H0,S0}
 *     group 2  RF strut        same                                                   11 rows, own {H1,S1}
 *     group 3  rear axle       5 x dball(chassis, axle)                               5 rows, own {A}
 * so A = blockdiag(D_g) + U M_C^-1 U^T with U the chassis columns of J.  Eliminating the groups first
 * (block LDL^T of each D_g, at most 11x11) leaves a 6x6 Schur system on the chassis acceleration z:
 *     (M_C + sum_g U_g^T D_g^-1 U_g) z = sum_g U_g^T D_g^-1 r_g ,   lambda_g = D_g^-1 (r_g - U_g z).
 * This is an exact (direct) solve of the same system -- no iteration count, no warm start -- at about a
 * quarter of the flops and a tenth of the working set of the dense factorisation.
 */
#pragma once
#include "pd_car.h"

namespace pd {

#define PD_GMAX 11

struct Sym3 { float xx, xy, xz, yy, yz, zz; };
PD_HD V3 sym3_mul(const Sym3& m, V3 v) { return v3(m.xx * v.x + m.xy * v.y + m.xz * v.z, m.xy * v.x + m.yy * v.y + m.yz * v.z, m.xz * v.x + m.yz * v.y + m.zz * v.z); }
/* R diag(d) R^T with R's columns = frame axes */
PD_HD Sym3 rot_diag(const Frame& f, V3 d) {
    Sym3 m;
    m.xx = f.ax.x * d.x * f.ax.x + f.ay.x * d.y * f.ay.x + f.az.x * d.z * f.az.x;
    m.xy = f.ax.x * d.x * f.ax.y + f.ay.x * d.y * f.ay.y + f.az.x * d.z * f.az.y;
    m.xz = f.ax.x * d.x * f.ax.z + f.ay.x * d.y * f.ay.z + f.az.x * d.z * f.az.z;
    m.yy = f.ax.y * d.x * f.ax.y + f.ay.y * d.y * f.ay.y + f.az.y * d.z * f.az.y;
    m.yz = f.ax.y * d.x * f.ax.z + f.ay.y * d.y * f.ay.z + f.az.y * d.z * f.az.z;
    m.zz = f.ax.z * d.x * f.ax.z + f.ay.z * d.y * f.ay.z + f.az.z * d.z * f.az.z;
    return m;
}

struct BodyDyn {      /* per-body quantities of dxStepIsland stage 0/1 */
    float invMass;
    Sym3 invI;        /* world-frame inverse inertia */
    float t1[6];      /* invM*f + v/h */
};

/* gyroscopic torque, implicit form (ODE step.cpp, dxBodyGyroscopic default) */
PD_HD V3 gyro_torque(const Body& b, const Sym3& I, float h) {
    const V3 L = sym3_mul(I, b.w);
    /* Itild = -[L]x * h + I */
    float m[9] = {I.xx, L.z * h + I.xy, -L.y * h + I.xz,
                  -L.z * h + I.xy, I.yy, L.x * h + I.yz,
                  L.y * h + I.xz, -L.x * h + I.yz, I.zz};
    const float hinv = 1.0f / h;
    const V3 Ls = v3(L.x * hinv, L.y * hinv, L.z * hinv);
    const float c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const float det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0) return v3(0, 0, 0);
    const float id = 1.0f / det;
    float inv[9];
    inv[0] = c00 * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c01 * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c02 * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    const float Ir[9] = {I.xx, I.xy, I.xz, I.xy, I.yy, I.yz, I.xz, I.yz, I.zz};
    float M[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r * 3 + c] = Ir[r * 3 + 0] * inv[0 * 3 + c] + Ir[r * 3 + 1] * inv[1 * 3 + c] + Ir[r * 3 + 2] * inv[2 * 3 + c];
    M[0] -= 1; M[4] -= 1; M[8] -= 1;
    return v3(M[0] * Ls.x + M[1] * Ls.y + M[2] * Ls.z, M[3] * Ls.x + M[4] * Ls.y + M[5] * Ls.z, M[6] * Ls.x + M[7] * Ls.y + M[8] * Ls.z);
}

/* ---- constraint rows (ode/src/joints/*.cpp getInfo2, see oracle/ode_restate/ode_core.h) ---- */
PD_HD void set6(float* J, V3 l, V3 a) { J[0] = l.x; J[1] = l.y; J[2] = l.z; J[3] = a.x; J[4] = a.y; J[5] = a.z; }

/* dball row: J0 for body 0, J1 for body 1 */
PD_HD void row_dball(const Body& b0, const Body& b1, V3 anchor1, V3 anchor2, float target, float k, float* J0, float* J1, float& c) {
    const V3 ra1 = rot(b0.fr, anchor1), ra2 = rot(b1.fr, anchor2);
    const V3 g1 = v3(ra1.x + b0.fr.p.x, ra1.y + b0.fr.p.y, ra1.z + b0.fr.p.z);
    const V3 g2 = v3(ra2.x + b1.fr.p.x, ra2.y + b1.fr.p.y, ra2.z + b1.fr.p.z);
    V3 q = g1 - g2;
    const float dist = sqrtf(dot(q, q));
    if (dist < 1e-7f) {
        q = body_point_vel(b0, g1) - body_point_vel(b1, g2);
        if (sqrtf(dot(q, q)) < 1e-7f) q = v3(1, 0, 0);
    }
    { float l = dot(q, q); if (l > 0) { l = 1.0f / sqrtf(l); q = q * l; } else q = v3(1, 0, 0); }
    set6(J0, q, cross(ra1, q));
    set6(J1, neg(q), cross(q, ra2));
    c = k * (target - dist);
}

/* setFixedOrientation: 3 angular rows; J on body0 = +I (angular), body1 = -I; returns c[3] */
PD_HD void fixed_orientation_c(const Body& b0, const Body& b1, Quat qrel, float k2, float* c3) {
    Quat qq = qmul1(b0.q, b1.q);
    Quat qerr = qmul2(qq, qrel);
    if (qerr.w < 0) { qerr.x = -qerr.x; qerr.y = -qerr.y; qerr.z = -qerr.z; }
    const V3 e = rot(b0.fr, v3(qerr.x, qerr.y, qerr.z));
    c3[0] = k2 * e.x; c3[1] = k2 * e.y; c3[2] = k2 * e.z;
}

PD_HD void plane_space(V3 n, V3& p, V3& q) {
    if (fabsf(n.z) > 0.70710678118654752440f) {
        const float a = n.y * n.y + n.z * n.z, k = 1.0f / sqrtf(a);
        p = v3(0, -n.z * k, n.y * k);
        q = v3(a * k, -n.x * p.z, n.x * p.y);
    } else {
        const float a = n.x * n.x + n.y * n.y, k = 1.0f / sqrtf(a);
        p = v3(-n.y * k, n.x * k, 0);
        q = v3(-n.z * p.y, n.z * p.x, a * k);
    }
}

/* Scratch of one joint group, addressed with a COMPILE-TIME stride so that the same code runs on
 *   - shared memory interleaved over the lanes of a block (GPU: element k of lane t at base[k * blockDim + t]:
 *     every lane owns one bank column -> conflict-free, and with all loops unrolled every access is an LDS/STS
 *     with an immediate offset), and
 *   - a plain local array (host debugging build, stride 1).
 * Every group is padded to PD_GMAX = 11 rows (pad rows: zero Jacobians, unit diagonal) so that the four lanes of
 * a quad run the same straight-line code whatever their group's real size (11 / 11 / 5 / 6 rows).
 * Layout (words): rows part JA 11x6 | JB 11x6;  D part Y 11x7 ([U | c], then [U | r], then L^-1[U | r]; column 6 ends
 * as lambda) | D packed lower 66 (in place: unit-lower L below the diagonal) | dg 11 (cfm per row, then the D diagonal). */
#define PD_GSCR_WORDS 286
#define PD_GSCR_ROWS_WORDS 132   /* JA | JB : read by the D build and by the final J^T lambda */
#define PD_GSCR_D_WORDS 154      /* Y | D | dg : what the factorisation and the substitutions work on */
/* S: stride of the JA | JB part at p;  SD: stride of the Y | D | dg part at q.  The two parts may live in different
 * memories (e.g. rows in local memory, stride 1, and the rest in shared memory, lane-interleaved). */
template <int S, int SD = S> struct GScr {
    float* p; float* q;
    PD_HD void bind(float* base) { p = base; q = base + PD_GSCR_ROWS_WORDS * S; }          /* one contiguous scratch */
    PD_HD void bind(float* rows, float* dpart) { p = rows; q = dpart; }
    PD_HD float& JA(int i, int k) const { return p[(i * 6 + k) * S]; }
    PD_HD float& JB(int i, int k) const { return p[(66 + i * 6 + k) * S]; }
    PD_HD float& Y(int i, int k) const { return q[(i * 7 + k) * SD]; }
    PD_HD float& D(int i, int j) const { return q[(77 + i * (i + 1) / 2 + j) * SD]; }   /* j <= i */
    PD_HD float& dg(int i) const { return q[(143 + i) * SD]; }
};
/* all rows zero; pad rows get cfm = h so that their diagonal becomes cfm/h = 1 */
template <class GS> PD_HD void zero_group(const GS& G, int n, float cfm, float h) {
    PD_NOUNROLL
    for (int i = 0; i < PD_GMAX; ++i) {
        PD_UNROLL
        for (int k = 0; k < 6; ++k) { G.JA(i, k) = 0; G.JB(i, k) = 0; G.Y(i, k) = 0; }
        G.Y(i, 6) = 0; G.dg(i) = (i < n) ? cfm : h;
    }
}
/* dball row between the chassis (body 0 -> U) and an own body (body 1 -> JA); c goes to Y(i,6) */
template <class GS> PD_HD void grow_dball(const GS& G, int i, const Body& b0, const Body& b1, V3 anchor1, V3 anchor2, float target, float k, float cfm) {
    float J0[6], J1[6], c;
    row_dball(b0, b1, anchor1, anchor2, target, k, J0, J1, c);
    PD_UNROLL
    for (int q = 0; q < 6; ++q) { G.Y(i, q) = J0[q]; G.JA(i, q) = J1[q]; }
    G.Y(i, 6) = c; G.dg(i) = cfm;
}

/* group "tank": fixed joint tank(b0) <-> chassis(b1)  (fixed.cpp getInfo2: rows 0-2 linear, 3-5 angular) */
template <class GS> PD_HD void build_tank(const PdCarParams& P, const Body& T, const Body& C, float fps, const GS& G) {
    zero_group(G, 6, P.worldCFM, 1.0f / fps);
    const V3 ofs = rot(T.fr, v3(P.tankOffset[0], P.tankOffset[1], P.tankOffset[2]));
    /* linear rows: J1l = I, J1a = [ofs]x (rows), J2l = -I */
    G.JA(0, 0) = 1; G.JA(1, 1) = 1; G.JA(2, 2) = 1;
    G.JA(0, 4) = -ofs.z; G.JA(0, 5) = ofs.y; G.JA(1, 3) = ofs.z; G.JA(1, 5) = -ofs.x; G.JA(2, 3) = -ofs.y; G.JA(2, 4) = ofs.x;
    G.Y(0, 0) = -1; G.Y(1, 1) = -1; G.Y(2, 2) = -1;
    const float k = fps * P.worldERP;
    G.Y(0, 6) = k * (C.fr.p.x - T.fr.p.x + ofs.x); G.Y(1, 6) = k * (C.fr.p.y - T.fr.p.y + ofs.y); G.Y(2, 6) = k * (C.fr.p.z - T.fr.p.z + ofs.z);
    G.JA(3, 3) = 1; G.JA(4, 4) = 1; G.JA(5, 5) = 1; G.Y(3, 3) = -1; G.Y(4, 4) = -1; G.Y(5, 5) = -1;
    Quat qrel; qrel.w = P.tankQrel[0]; qrel.x = P.tankQrel[1]; qrel.y = P.tankQrel[2]; qrel.z = P.tankQrel[3];
    float c3[3]; fixed_orientation_c(T, C, qrel, fps * P.worldERP * 2.0f, c3);
    G.Y(3, 6) = c3[0]; G.Y(4, 6) = c3[1]; G.Y(5, 6) = c3[2];
}

/* groups "strut": A = hub, B = strut body */
template <class GS> PD_HD void build_strut(const PdCarParams& P, const PdStrut& S, const Body& C, const Body& H, const Body& B, V3 steerA1, V3 steerA2, float fps, float dballErp, float dballCfm, const GS& G) {
    zero_group(G, 11, P.worldCFM, 1.0f / fps);
    PD_UNROLL
    for (int l = 0; l < 3; ++l) {
        V3 a1 = v3(S.link[l].anchor1[0], S.link[l].anchor1[1], S.link[l].anchor1[2]);
        V3 a2 = v3(S.link[l].anchor2[0], S.link[l].anchor2[1], S.link[l].anchor2[2]);
        if (l == 2) { a1 = steerA1; a2 = steerA2; }
        grow_dball(G, l, C, H, a1, a2, S.link[l].distance, fps * dballErp, dballCfm);
    }
    { /* slider (b0 = strut body, b1 = hub): rows 3..7 */
        Quat qrel; qrel.w = S.sliderQrel[0]; qrel.x = S.sliderQrel[1]; qrel.y = S.sliderQrel[2]; qrel.z = S.sliderQrel[3];
        G.JB(3, 3) = 1; G.JB(4, 4) = 1; G.JB(5, 5) = 1; G.JA(3, 3) = -1; G.JA(4, 4) = -1; G.JA(5, 5) = -1;
        float c3[3]; fixed_orientation_c(B, H, qrel, fps * P.worldERP * 2.0f, c3);
        G.Y(3, 6) = c3[0]; G.Y(4, 6) = c3[1]; G.Y(5, 6) = c3[2];
        V3 c = H.fr.p - B.fr.p;
        const V3 ax1 = rot(B.fr, v3(S.sliderAxis1[0], S.sliderAxis1[1], S.sliderAxis1[2]));
        V3 p, q; plane_space(ax1, p, q);
        const V3 cp = cross(c, p) * 0.5f, cq = cross(c, q) * 0.5f;
        G.JB(6, 0) = p.x; G.JB(6, 1) = p.y; G.JB(6, 2) = p.z; G.JB(6, 3) = cp.x; G.JB(6, 4) = cp.y; G.JB(6, 5) = cp.z;
        G.JA(6, 0) = -p.x; G.JA(6, 1) = -p.y; G.JA(6, 2) = -p.z; G.JA(6, 3) = cp.x; G.JA(6, 4) = cp.y; G.JA(6, 5) = cp.z;
        G.JB(7, 0) = q.x; G.JB(7, 1) = q.y; G.JB(7, 2) = q.z; G.JB(7, 3) = cq.x; G.JB(7, 4) = cq.y; G.JB(7, 5) = cq.z;
        G.JA(7, 0) = -q.x; G.JA(7, 1) = -q.y; G.JA(7, 2) = -q.z; G.JA(7, 3) = cq.x; G.JA(7, 4) = cq.y; G.JA(7, 5) = cq.z;
        const V3 ofs = rot(H.fr, v3(S.sliderOffset[0], S.sliderOffset[1], S.sliderOffset[2]));
        c = c + ofs;
        const float k = fps * P.worldERP;
        G.Y(6, 6) = k * dot(p, c); G.Y(7, 6) = k * dot(q, c);
    }
    { /* ball (b0 = chassis, b1 = strut body): rows 8..10 */
        const V3 a1 = rot(C.fr, v3(S.ballAnchor1[0], S.ballAnchor1[1], S.ballAnchor1[2]));
        const V3 a2 = rot(B.fr, v3(S.ballAnchor2[0], S.ballAnchor2[1], S.ballAnchor2[2]));
        G.Y(8, 0) = 1; G.Y(9, 1) = 1; G.Y(10, 2) = 1; G.JB(8, 0) = -1; G.JB(9, 1) = -1; G.JB(10, 2) = -1;
        G.Y(8, 4) = a1.z; G.Y(8, 5) = -a1.y; G.Y(9, 3) = -a1.z; G.Y(9, 5) = a1.x; G.Y(10, 3) = a1.y; G.Y(10, 4) = -a1.x;
        G.JB(8, 4) = -a2.z; G.JB(8, 5) = a2.y; G.JB(9, 3) = a2.z; G.JB(9, 5) = -a2.x; G.JB(10, 3) = -a2.y; G.JB(10, 4) = a2.x;
        const float k = fps * P.worldERP;
        G.Y(8, 6) = k * (a2.x + B.fr.p.x - a1.x - C.fr.p.x); G.Y(9, 6) = k * (a2.y + B.fr.p.y - a1.y - C.fr.p.y); G.Y(10, 6) = k * (a2.z + B.fr.p.z - a1.z - C.fr.p.z);
    }
}

/* group "axle": dball links (b0 = chassis, b1 = axle) */
template <class GS> PD_HD void build_axle(const PdCarParams& P, const Body& C, const Body& A, float fps, float dballErp, float dballCfm, const GS& G) {
    const int n = P.axle.nLinks;
    zero_group(G, n, dballCfm, 1.0f / fps);
    PD_UNROLL
    for (int l = 0; l < PD_AXLE_LINKS; ++l) {
        if (l < n) {
            const PdDBall& K = P.axle.link[l];
            grow_dball(G, l, C, A, v3(K.anchor1[0], K.anchor1[1], K.anchor1[2]), v3(K.anchor2[0], K.anchor2[1], K.anchor2[2]), K.distance, fps * dballErp, dballCfm);
        }
    }
}

PD_HD void jinvm6(const float* J, const BodyDyn& d, float* o) {
    o[0] = J[0] * d.invMass; o[1] = J[1] * d.invMass; o[2] = J[2] * d.invMass;
    const V3 a = sym3_mul(d.invI, v3(J[3], J[4], J[5]));
    o[3] = a.x; o[4] = a.y; o[5] = a.z;
}

/* factor one group: D = JA MA^-1 JA^T + JB MB^-1 JB^T + cfm/h ; L D L^T in place ; Y <- L^-1 [U | r];
 * accumulates the chassis Schur complement S (6x6 lower, packed 21) and right-hand side b6.
 * Row loops are kept ROLLED (compact code: the kernel is instruction-fetch sensitive), the 6- and 7-wide inner
 * loops are unrolled.  n = rows to process: the 4-lanes-per-car kernel runs all PD_GMAX rows on every lane (padding
 * rows are identity, so there is no divergence); the thread-per-car kernel passes the group's real row count. */
/* D = JA MA^-1 JA^T + JB MB^-1 JB^T + cfm/h and the right-hand side r, row pairs first, first + step, ... (step = 2 when a helper
 * lane takes every other pair: the two callers write disjoint entries of the shared scratch and must synchronise afterwards) */
template <class GS> PD_HDN void build_D(const GS& G, const BodyDyn& dA, const BodyDyn& dB, const BodyDyn& dC, float hinv, const int n = PD_GMAX, const bool hasB = true, const int first = 0, const int step = 1) {
    /* D build, TWO rows per pass: rows i and i+1 share every load of the rows j <= i they are multiplied with
     * (half the scratch reads, two independent sums in flight); each sum runs over the same terms in the same order
     * as a row-at-a-time build, so the result is bit-identical to it. */
    PD_NOUNROLL
    for (int i = 2 * first; i < n; i += 2 * step) {
        const bool two = i + 1 < n;
        const int i1 = two ? i + 1 : i;
        float ra0[6], rb0[6], ja0[6], jb0[6], ra1[6], rb1[6], ja1[6], jb1[6];
        PD_UNROLL
        for (int k = 0; k < 6; ++k) { ra0[k] = G.JA(i, k); rb0[k] = G.JB(i, k); ra1[k] = G.JA(i1, k); rb1[k] = G.JB(i1, k); }
        jinvm6(ra0, dA, ja0); jinvm6(ra1, dA, ja1);
        if (hasB) { jinvm6(rb0, dB, jb0); jinvm6(rb1, dB, jb1); }
        PD_NOUNROLL
        for (int j = 0; j <= i; ++j) {
            float a[6], b[6];
            PD_UNROLL
            for (int k = 0; k < 6; ++k) { a[k] = G.JA(j, k); if (hasB) b[k] = G.JB(j, k); }
            float s0 = ja0[0] * a[0] + ja0[1] * a[1] + ja0[2] * a[2] + ja0[3] * a[3] + ja0[4] * a[4] + ja0[5] * a[5];
            float s1 = ja1[0] * a[0] + ja1[1] * a[1] + ja1[2] * a[2] + ja1[3] * a[3] + ja1[4] * a[4] + ja1[5] * a[5];
            if (hasB) {
                s0 += jb0[0] * b[0] + jb0[1] * b[1] + jb0[2] * b[2] + jb0[3] * b[3] + jb0[4] * b[4] + jb0[5] * b[5];
                s1 += jb1[0] * b[0] + jb1[1] * b[1] + jb1[2] * b[2] + jb1[3] * b[3] + jb1[4] * b[4] + jb1[5] * b[5];
            }
            G.D(i, j) = s0;
            if (two) G.D(i1, j) = s1;
        }
        G.D(i, i) += G.dg(i) * hinv;
        if (two) {      /* the second row's own diagonal entry */
            float s1 = ja1[0] * ra1[0] + ja1[1] * ra1[1] + ja1[2] * ra1[2] + ja1[3] * ra1[3] + ja1[4] * ra1[4] + ja1[5] * ra1[5];
            if (hasB) s1 += jb1[0] * rb1[0] + jb1[1] * rb1[1] + jb1[2] * rb1[2] + jb1[3] * rb1[3] + jb1[4] * rb1[4] + jb1[5] * rb1[5];
            G.D(i1, i1) = s1;
            G.D(i1, i1) += G.dg(i1) * hinv;
        }
        /* r_i = c_i/h - J_i (v/h + M^-1 f) */
        {
            float s = ra0[0] * dA.t1[0] + ra0[1] * dA.t1[1] + ra0[2] * dA.t1[2] + ra0[3] * dA.t1[3] + ra0[4] * dA.t1[4] + ra0[5] * dA.t1[5];
            s += G.Y(i, 0) * dC.t1[0] + G.Y(i, 1) * dC.t1[1] + G.Y(i, 2) * dC.t1[2] + G.Y(i, 3) * dC.t1[3] + G.Y(i, 4) * dC.t1[4] + G.Y(i, 5) * dC.t1[5];
            if (hasB) s += rb0[0] * dB.t1[0] + rb0[1] * dB.t1[1] + rb0[2] * dB.t1[2] + rb0[3] * dB.t1[3] + rb0[4] * dB.t1[4] + rb0[5] * dB.t1[5];
            G.Y(i, 6) = G.Y(i, 6) * hinv - s;
        }
        if (two) {
            float s = ra1[0] * dA.t1[0] + ra1[1] * dA.t1[1] + ra1[2] * dA.t1[2] + ra1[3] * dA.t1[3] + ra1[4] * dA.t1[4] + ra1[5] * dA.t1[5];
            s += G.Y(i1, 0) * dC.t1[0] + G.Y(i1, 1) * dC.t1[1] + G.Y(i1, 2) * dC.t1[2] + G.Y(i1, 3) * dC.t1[3] + G.Y(i1, 4) * dC.t1[4] + G.Y(i1, 5) * dC.t1[5];
            if (hasB) s += rb1[0] * dB.t1[0] + rb1[1] * dB.t1[1] + rb1[2] * dB.t1[2] + rb1[3] * dB.t1[3] + rb1[4] * dB.t1[4] + rb1[5] * dB.t1[5];
            G.Y(i1, 6) = G.Y(i1, 6) * hinv - s;
        }
    }
}

template <class GS> PD_HDN void factor_group(const GS& G, const BodyDyn& dA, const BodyDyn& dB, const BodyDyn& dC, float hinv, float* S21, float* b6, const int n = PD_GMAX, const bool hasB = true, const bool dBuilt = false) {
    if (!dBuilt) build_D(G, dA, dB, dC, hinv, n, hasB);
    /* L D L^T, row by row (same recurrence as the oracle's dense factorisation), in place */
    PD_NOUNROLL
    for (int i = 0; i < n; ++i) {
        PD_NOUNROLL
        for (int j = 0; j < i; ++j) {
            float s = G.D(i, j);
            PD_UNROLL4
            for (int k = 0; k < j; ++k) s -= G.D(i, k) * G.D(j, k);    /* D(i,k) = u_k (unscaled), D(j,k) = L_jk */
            G.D(i, j) = s;
        }
        float dii = G.D(i, i);
        PD_UNROLL4
        for (int j = 0; j < i; ++j) { const float u = G.D(i, j); const float lij = u / G.dg(j); dii -= u * lij; G.D(i, j) = lij; }
        G.dg(i) = dii;
    }
    /* forward substitution on the 7 right-hand sides, two rows per pass (shared loads of the rows above, same
     * subtraction order per row as the row-at-a-time form: bit-identical) */
    PD_NOUNROLL
    for (int i = 0; i < n; i += 2) {
        const bool two = i + 1 < n;
        const int i1 = two ? i + 1 : i;
        float y0[7], y1[7];
        PD_UNROLL
        for (int k = 0; k < 7; ++k) { y0[k] = G.Y(i, k); y1[k] = G.Y(i1, k); }
        PD_NOUNROLL
        for (int j = 0; j < i; ++j) {
            const float l0 = G.D(i, j), l1 = G.D(i1, j);
            PD_UNROLL
            for (int k = 0; k < 7; ++k) { const float yj = G.Y(j, k); y0[k] -= l0 * yj; y1[k] -= l1 * yj; }
        }
        PD_UNROLL
        for (int k = 0; k < 7; ++k) G.Y(i, k) = y0[k];
        /* S += Yu^T D^-1 Yu ; b += Yu^T D^-1 yr */
        {
            const float di = 1.0f / G.dg(i);
            PD_UNROLL
            for (int a = 0; a < 6; ++a) {
                const float ya = y0[a] * di;
                PD_UNROLL
                for (int bb = 0; bb < 6; ++bb) if (bb <= a) S21[a * (a + 1) / 2 + bb] += ya * y0[bb];
                b6[a] += ya * y0[6];
            }
        }
        if (two) {
            const float l = G.D(i1, i);
            PD_UNROLL
            for (int k = 0; k < 7; ++k) { y1[k] -= l * y0[k]; G.Y(i1, k) = y1[k]; }
            const float di = 1.0f / G.dg(i1);
            PD_UNROLL
            for (int a = 0; a < 6; ++a) {
                const float ya = y1[a] * di;
                PD_UNROLL
                for (int bb = 0; bb < 6; ++bb) if (bb <= a) S21[a * (a + 1) / 2 + bb] += ya * y1[bb];
                b6[a] += ya * y1[6];
            }
        }
    }
}

/* lambda_g = L^-T D^-1 (yr - Yu z);  cforce on own bodies = J^T lambda */
template <class GS> PD_HDN void backsolve_group(const GS& G, const float* z, float* cfA, float* cfB, const int n = PD_GMAX) {
    PD_NOUNROLL
    for (int i = 0; i < n; ++i) {
        float s = G.Y(i, 6);
        PD_UNROLL
        for (int k = 0; k < 6; ++k) s -= G.Y(i, k) * z[k];
        G.Y(i, 6) = s / G.dg(i);
    }
    PD_NOUNROLL
    for (int i = n - 1; i >= 0; --i) { float s = G.Y(i, 6); PD_UNROLL4 for (int k = i + 1; k < n; ++k) s -= G.D(k, i) * G.Y(k, 6); G.Y(i, 6) = s; }
    PD_UNROLL
    for (int k = 0; k < 6; ++k) { cfA[k] = 0; cfB[k] = 0; }
    PD_NOUNROLL
    for (int i = 0; i < n; ++i) {
        const float lam = G.Y(i, 6);
        PD_UNROLL
        for (int k = 0; k < 6; ++k) { cfA[k] += G.JA(i, k) * lam; cfB[k] += G.JB(i, k) * lam; }
    }
}

/* Serial variant of the back-substitution: fold the group into an affine map of the chassis unknown z,
 *     cforce_A = pA - QA z ,  cforce_B = pB - QB z      (W = L^-T D^-1 [Yu | yr];  p = J^T W[:,6],  Q = J^T W[:,0:6])
 * so that the group's scratch can be reused by the next group before z is known. */
template <class GS> PD_HDN void fold_group(const GS& G, float* pA, float* QA, float* pB, float* QB, const int n = PD_GMAX, const bool hasB = true) {
    PD_NOUNROLL
    for (int i = 0; i < n; ++i) { const float di = 1.0f / G.dg(i); PD_UNROLL for (int k = 0; k < 7; ++k) G.Y(i, k) *= di; }
    PD_NOUNROLL
    for (int i = n - 1; i >= 0; --i) {
        float y[7];
        PD_UNROLL
        for (int k = 0; k < 7; ++k) y[k] = G.Y(i, k);
        PD_NOUNROLL
        for (int r = i + 1; r < n; ++r) { const float l = G.D(r, i); PD_UNROLL for (int k = 0; k < 7; ++k) y[k] -= l * G.Y(r, k); }
        PD_UNROLL
        for (int k = 0; k < 7; ++k) G.Y(i, k) = y[k];
    }
    /* p = J^T W[:,6], Q = J^T W[:,0:6]: one pass over the rows with the 2 x 42 sums kept in registers
     * (each sum still runs over i ascending from 0, as the row-by-row form did) */
    PD_NOUNROLL
    for (int body = 0; body < 2; ++body) {
        float* pOut = body ? pB : pA; float* QOut = body ? QB : QA;
        float acc[42];
        PD_UNROLL
        for (int k = 0; k < 42; ++k) acc[k] = 0;
        if (body == 0 || hasB) {
            PD_NOUNROLL
            for (int i = 0; i < n; ++i) {
                float y[7], jr[6];
                PD_UNROLL
                for (int k = 0; k < 7; ++k) y[k] = G.Y(i, k);
                PD_UNROLL
                for (int a = 0; a < 6; ++a) jr[a] = body ? G.JB(i, a) : G.JA(i, a);
                PD_UNROLL
                for (int a = 0; a < 6; ++a) {
                    PD_UNROLL
                    for (int k = 0; k < 7; ++k) acc[a * 7 + k] += jr[a] * y[k];
                }
            }
        }
        PD_UNROLL
        for (int a = 0; a < 6; ++a) {
            pOut[a] = acc[a * 7 + 6];
            PD_UNROLL
            for (int k = 0; k < 6; ++k) QOut[a * 6 + k] = acc[a * 7 + k];
        }
    }
}
PD_HD void unfold(const float* p, const float* Q, const float* z, float* cf) {
    PD_UNROLL
    for (int a = 0; a < 6; ++a) { float s = p[a]; PD_UNROLL for (int k = 0; k < 6; ++k) s -= Q[a * 6 + k] * z[k]; cf[a] = s; }
}

/* util.cpp dxStepBody (finite rotation mode 1, no axis) */
PD_HD void integrate_body(Body& b, float h) {
    b.fr.p.x += h * b.v.x; b.fr.p.y += h * b.v.y; b.fr.p.z += h * b.v.z;
    const float wlen = sqrtf(b.w.x * b.w.x + b.w.y * b.w.y + b.w.z * b.w.z);
    h *= 0.5f;
    const float theta = wlen * h;
    Quat q; q.w = m_cos(theta);
    const float sinc = (fabsf(theta) < 1.0e-4f) ? 1.0f - theta * theta * 0.166666666666666666667f : m_sin(theta) / theta;
    const float s = sinc * h;
    q.x = b.w.x * s; q.y = b.w.y * s; q.z = b.w.z * s;
    Quat q2 = qmul0(q, b.q);
    float l = q2.w * q2.w + q2.x * q2.x + q2.y * q2.y + q2.z * q2.z;
    if (l > 0) { l = 1.0f / sqrtf(l); q2.w *= l; q2.x *= l; q2.y *= l; q2.z *= l; } else { q2.w = 1; q2.x = q2.y = q2.z = 0; }
    b.q = q2;
    quat_to_axes(b.q, b.fr.ax, b.fr.ay, b.fr.az);
}

PD_HD void apply_update(Body& b, const BodyDyn& d, const float* cf, float h) {
    const float imh = h * d.invMass;
    b.v.x += (cf[0] + b.F.x) * imh; b.v.y += (cf[1] + b.F.y) * imh; b.v.z += (cf[2] + b.F.z) * imh;
    const V3 t = v3((cf[3] + b.T.x) * h, (cf[4] + b.T.y) * h, (cf[5] + b.T.z) * h);
    const V3 dw = sym3_mul(d.invI, t);
    b.w += dw;
}

/* dxStepIsland stage 0/1 for one body: world inertia, gyroscopic torque, gravity, invM*f + v/h */
PD_HD void body_dyn(Body& b, float gravityY, float h, BodyDyn& d) {
    const float hinv = 1.0f / h;
    const Sym3 I = rot_diag(b.fr, b.I);
    d.invI = rot_diag(b.fr, v3(1.0f / b.I.x, 1.0f / b.I.y, 1.0f / b.I.z));
    d.invMass = 1.0f / b.mass;
    b.T += gyro_torque(b, I, h);
    b.F.y += b.mass * gravityY;
    d.t1[0] = b.F.x * d.invMass + b.v.x * hinv; d.t1[1] = b.F.y * d.invMass + b.v.y * hinv; d.t1[2] = b.F.z * d.invMass + b.v.z * hinv;
    const V3 a = sym3_mul(d.invI, b.T);
    d.t1[3] = a.x + b.w.x * hinv; d.t1[4] = a.y + b.w.y * hinv; d.t1[5] = a.z + b.w.z * hinv;
}
/* S += M_C (mass on the linear diagonal, world inertia on the angular block) */
PD_HD void schur_add_chassis(float* S21, const Body& C) {
    const Sym3 Ic = rot_diag(C.fr, C.I);
    S21[0] += C.mass; S21[2] += C.mass; S21[5] += C.mass;
    S21[9] += Ic.xx; S21[13] += Ic.xy; S21[14] += Ic.yy; S21[18] += Ic.xz; S21[19] += Ic.yz; S21[20] += Ic.zz;
}
/* 6x6 LDL^T solve S z = b (S packed lower, row-major) */
PD_HD void solve6(const float* S21, const float* b6, float* z) {
    float Lm[6][6], d[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 6; ++j) {
            if (j < i) { float s = S21[i * (i + 1) / 2 + j]; for (int k = 0; k < 6; ++k) if (k < j) s -= Lm[i][k] * Lm[j][k] * d[k]; Lm[i][j] = s / d[j]; }
        }
        float s = S21[i * (i + 1) / 2 + i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 6; ++k) if (k < i) s -= Lm[i][k] * Lm[i][k] * d[k];
        d[i] = s;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) { float s = b6[i]; for (int k = 0; k < 6; ++k) if (k < i) s -= Lm[i][k] * z[k]; z[i] = s; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) z[i] /= d[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 5; i >= 0; --i) { float s = z[i]; for (int k = 0; k < 6; ++k) if (k > i) s -= Lm[k][i] * z[k]; z[i] = s; }
}
/* chassis: z = M_C^-1 U^T lambda  ->  dv = h (M_C^-1 f + z) */
PD_HD void chassis_update(Body& Cb, const BodyDyn& d, const float* z, float h) {
    const float imh = h * d.invMass;
    Cb.v.x += Cb.F.x * imh + h * z[0]; Cb.v.y += Cb.F.y * imh + h * z[1]; Cb.v.z += Cb.F.z * imh + h * z[2];
    const V3 dw = sym3_mul(d.invI, v3(Cb.T.x * h, Cb.T.y * h, Cb.T.z * h));
    Cb.w.x += dw.x + h * z[3]; Cb.w.y += dw.y + h * z[4]; Cb.w.z += dw.z + h * z[5];
}

/* dWorldStep for the car's island, one thread doing the four groups one after the other with ONE scratch
 * (thread-per-car kernel for large batches, and the host debugging build) */
template <int STRIDE, int STRIDE_D> PD_HDN void world_step(const PdCarParams& P, Body* b, const V3* steerAnchor1, const V3* steerAnchor2, float dballErp, float dballCfm, float h, float* scratch, float* scratchD) {
    const float hinv = 1.0f / h;
    BodyDyn dyn[7 /* demo topology only */];
    for (int i = 0; i < 7 /* demo topology only */; ++i) body_dyn(b[i], P.gravityY, h, dyn[i]);
    float S21[21], b6[6];
    for (int k = 0; k < 21; ++k) S21[k] = 0;
    for (int k = 0; k < 6; ++k) b6[k] = 0;
    GScr<STRIDE, STRIDE_D> G; G.bind(scratch, scratchD);   /* rows part: PD_GSCR_ROWS_WORDS words at STRIDE; D part: PD_GSCR_D_WORDS at STRIDE_D */
    /* folded groups: [tank | hub0, strut0 | hub1, strut1 | axle] */
    float pv[6][6], Qv[6][36], pdump[6], Qdump[36];
    const Body& C = b[PD_BODY_CHASSIS];
    build_tank(P, b[PD_BODY_TANK], C, hinv, G);
    factor_group(G, dyn[PD_BODY_TANK], dyn[PD_BODY_TANK], dyn[PD_BODY_CHASSIS], hinv, S21, b6, 6, false);
    fold_group(G, pv[0], Qv[0], pdump, Qdump, 6, false);
    for (int s = 0; s < 2; ++s) {
        build_strut(P, P.strut[s], C, b[PD_BODY_HUB0 + 2 * s], b[PD_BODY_STRUT0 + 2 * s], steerAnchor1[s], steerAnchor2[s], hinv, dballErp, dballCfm, G);
        factor_group(G, dyn[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_STRUT0 + 2 * s], dyn[PD_BODY_CHASSIS], hinv, S21, b6);
        fold_group(G, pv[1 + 2 * s], Qv[1 + 2 * s], pv[2 + 2 * s], Qv[2 + 2 * s]);
    }
    build_axle(P, C, b[PD_BODY_AXLE], hinv, dballErp, dballCfm, G);
    factor_group(G, dyn[PD_BODY_AXLE], dyn[PD_BODY_AXLE], dyn[PD_BODY_CHASSIS], hinv, S21, b6, 5, false);
    fold_group(G, pv[5], Qv[5], pdump, Qdump, 5, false);
    schur_add_chassis(S21, C);
    float z[6];
    solve6(S21, b6, z);
    const int own[6] = {PD_BODY_TANK, PD_BODY_HUB0, PD_BODY_STRUT0, PD_BODY_HUB1, PD_BODY_STRUT1, PD_BODY_AXLE};
    for (int g = 0; g < 6; ++g) { float cf[6]; unfold(pv[g], Qv[g], z, cf); apply_update(b[own[g]], dyn[own[g]], cf, h); }
    chassis_update(b[PD_BODY_CHASSIS], dyn[PD_BODY_CHASSIS], z, h);
    for (int i = 0; i < 7 /* demo topology only */; ++i) { integrate_body(b[i], h); b[i].F = v3(0, 0, 0); b[i].T = v3(0, 0, 0); }
}

} // namespace pd
