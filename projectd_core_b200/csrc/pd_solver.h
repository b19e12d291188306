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

struct GroupSys {
    int n, hasB;
    float JA[PD_GMAX][6], JB[PD_GMAX][6], U[PD_GMAX][6];
    float c[PD_GMAX], cfm[PD_GMAX];
};
struct GroupFac {
    float L[PD_GMAX * (PD_GMAX - 1) / 2];   /* strict lower triangle of the unit-lower factor, row-major packed */
    float d[PD_GMAX];                         /* diagonal of D */
    float Y[PD_GMAX][7];                      /* L^-1 [U | r] */
};
PD_HD int tri(int i, int j) { return i * (i - 1) / 2 + j; }   /* j < i */

PD_HD void zero_group(GroupSys& G, int n, int hasB, float cfm) {
    G.n = n; G.hasB = hasB;
    for (int i = 0; i < n; ++i) { for (int k = 0; k < 6; ++k) { G.JA[i][k] = 0; G.JB[i][k] = 0; G.U[i][k] = 0; } G.c[i] = 0; G.cfm[i] = cfm; }
}

/* group 0: fixed joint tank(b0) <-> chassis(b1)  (fixed.cpp getInfo2: rows 0-2 linear, 3-5 angular) */
PD_HD void build_tank(const PdCarParams& P, const Body& T, const Body& C, float fps, GroupSys& G) {
    zero_group(G, 6, 0, P.worldCFM);
    const V3 ofs = rot(T.fr, v3(P.tankOffset[0], P.tankOffset[1], P.tankOffset[2]));
    /* linear rows: J1l = I, J1a = [ofs]x (rows), J2l = -I */
    G.JA[0][0] = 1; G.JA[1][1] = 1; G.JA[2][2] = 1;
    G.JA[0][4] = -ofs.z; G.JA[0][5] = ofs.y; G.JA[1][3] = ofs.z; G.JA[1][5] = -ofs.x; G.JA[2][3] = -ofs.y; G.JA[2][4] = ofs.x;
    G.U[0][0] = -1; G.U[1][1] = -1; G.U[2][2] = -1;
    const float k = fps * P.worldERP;
    G.c[0] = k * (C.fr.p.x - T.fr.p.x + ofs.x); G.c[1] = k * (C.fr.p.y - T.fr.p.y + ofs.y); G.c[2] = k * (C.fr.p.z - T.fr.p.z + ofs.z);
    for (int i = 0; i < 3; ++i) { G.JA[3 + i][3 + i] = 1; G.U[3 + i][3 + i] = -1; }
    Quat qrel; qrel.w = P.tankQrel[0]; qrel.x = P.tankQrel[1]; qrel.y = P.tankQrel[2]; qrel.z = P.tankQrel[3];
    fixed_orientation_c(T, C, qrel, fps * P.worldERP * 2.0f, &G.c[3]);
}

/* groups 1,2: strut.  A = hub, B = strut body */
PD_HD void build_strut(const PdCarParams& P, const PdStrut& S, const Body& C, const Body& H, const Body& B, V3 steerA1, V3 steerA2, float fps, float dballErp, float dballCfm, GroupSys& G) {
    zero_group(G, 11, 1, P.worldCFM);
    for (int l = 0; l < 3; ++l) {
        V3 a1 = v3(S.link[l].anchor1[0], S.link[l].anchor1[1], S.link[l].anchor1[2]);
        V3 a2 = v3(S.link[l].anchor2[0], S.link[l].anchor2[1], S.link[l].anchor2[2]);
        if (l == 2) { a1 = steerA1; a2 = steerA2; }
        row_dball(C, H, a1, a2, S.link[l].distance, fps * dballErp, G.U[l], G.JA[l], G.c[l]);
        G.cfm[l] = dballCfm;
    }
    { /* slider (b0 = strut body, b1 = hub): rows 3..7 */
        Quat qrel; qrel.w = S.sliderQrel[0]; qrel.x = S.sliderQrel[1]; qrel.y = S.sliderQrel[2]; qrel.z = S.sliderQrel[3];
        for (int i = 0; i < 3; ++i) { G.JB[3 + i][3 + i] = 1; G.JA[3 + i][3 + i] = -1; }
        fixed_orientation_c(B, H, qrel, fps * P.worldERP * 2.0f, &G.c[3]);
        V3 c = H.fr.p - B.fr.p;
        const V3 ax1 = rot(B.fr, v3(S.sliderAxis1[0], S.sliderAxis1[1], S.sliderAxis1[2]));
        V3 p, q; plane_space(ax1, p, q);
        const V3 cp = cross(c, p) * 0.5f, cq = cross(c, q) * 0.5f;
        set6(G.JB[6], p, cp); set6(G.JA[6], neg(p), cp);
        set6(G.JB[7], q, cq); set6(G.JA[7], neg(q), cq);
        const V3 ofs = rot(H.fr, v3(S.sliderOffset[0], S.sliderOffset[1], S.sliderOffset[2]));
        c = c + ofs;
        const float k = fps * P.worldERP;
        G.c[6] = k * dot(p, c); G.c[7] = k * dot(q, c);
    }
    { /* ball (b0 = chassis, b1 = strut body): rows 8..10 */
        const V3 a1 = rot(C.fr, v3(S.ballAnchor1[0], S.ballAnchor1[1], S.ballAnchor1[2]));
        const V3 a2 = rot(B.fr, v3(S.ballAnchor2[0], S.ballAnchor2[1], S.ballAnchor2[2]));
        for (int i = 0; i < 3; ++i) { G.U[8 + i][i] = 1; G.JB[8 + i][i] = -1; }
        G.U[8][4] = a1.z; G.U[8][5] = -a1.y; G.U[9][3] = -a1.z; G.U[9][5] = a1.x; G.U[10][3] = a1.y; G.U[10][4] = -a1.x;
        G.JB[8][4] = -a2.z; G.JB[8][5] = a2.y; G.JB[9][3] = a2.z; G.JB[9][5] = -a2.x; G.JB[10][3] = -a2.y; G.JB[10][4] = a2.x;
        const float k = fps * P.worldERP;
        G.c[8] = k * (a2.x + B.fr.p.x - a1.x - C.fr.p.x); G.c[9] = k * (a2.y + B.fr.p.y - a1.y - C.fr.p.y); G.c[10] = k * (a2.z + B.fr.p.z - a1.z - C.fr.p.z);
    }
}

/* group 3: axle, 5 dball links (b0 = chassis, b1 = axle) */
PD_HD void build_axle(const PdCarParams& P, const Body& C, const Body& A, float fps, float dballErp, float dballCfm, GroupSys& G) {
    const int n = P.axle.nLinks;
    zero_group(G, n, 0, dballCfm);
    for (int l = 0; l < n; ++l) {
        const PdDBall& K = P.axle.link[l];
        row_dball(C, A, v3(K.anchor1[0], K.anchor1[1], K.anchor1[2]), v3(K.anchor2[0], K.anchor2[1], K.anchor2[2]), K.distance, fps * dballErp, G.U[l], G.JA[l], G.c[l]);
    }
}

PD_HD float dot6(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }
PD_HD void jinvm(const float* J, const BodyDyn& d, float* o) {
    o[0] = J[0] * d.invMass; o[1] = J[1] * d.invMass; o[2] = J[2] * d.invMass;
    const V3 a = sym3_mul(d.invI, v3(J[3], J[4], J[5]));
    o[3] = a.x; o[4] = a.y; o[5] = a.z;
}

/* factor one group: D = JA MA^-1 JA^T (+ JB MB^-1 JB^T) + cfm/h ; L D L^T ; Y = L^-1 [U | r];
 * accumulates the chassis Schur complement S (6x6 lower, packed 21) and right-hand side b6. */
PD_HDN void factor_group(const GroupSys& G, const BodyDyn& dA, const BodyDyn& dB, const BodyDyn& dC, float hinv, GroupFac& F, float* S21, float* b6) {
    const int n = G.n;
    float Dm[PD_GMAX][PD_GMAX];
    for (int i = 0; i < n; ++i) {
        float ja[6], jb[6];
        jinvm(G.JA[i], dA, ja);
        if (G.hasB) jinvm(G.JB[i], dB, jb);
        for (int j = 0; j <= i; ++j) {
            float s = dot6(ja, G.JA[j]);
            if (G.hasB) s += dot6(jb, G.JB[j]);
            Dm[i][j] = s;
        }
        Dm[i][i] += G.cfm[i] * hinv;
        /* r_i = c_i/h - J_i (v/h + M^-1 f) */
        float s = dot6(G.JA[i], dA.t1) + dot6(G.U[i], dC.t1);
        if (G.hasB) s += dot6(G.JB[i], dB.t1);
        F.Y[i][6] = G.c[i] * hinv - s;
        for (int k = 0; k < 6; ++k) F.Y[i][k] = G.U[i][k];
    }
    /* L D L^T, row by row (same recurrence as the oracle's dense factorisation) */
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < i; ++j) {
            float s = Dm[i][j];
            for (int k = 0; k < j; ++k) s -= Dm[i][k] * F.L[tri(j, k)];
            Dm[i][j] = s;
        }
        float dii = Dm[i][i];
        for (int j = 0; j < i; ++j) { const float lij = Dm[i][j] / F.d[j]; dii -= Dm[i][j] * lij; F.L[tri(i, j)] = lij; }
        F.d[i] = dii;
    }
    /* forward substitution on 7 right-hand sides */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) { const float l = F.L[tri(i, j)]; for (int k = 0; k < 7; ++k) F.Y[i][k] -= l * F.Y[j][k]; }
    /* S += Yu^T D^-1 Yu ; b += Yu^T D^-1 yr */
    for (int i = 0; i < n; ++i) {
        const float di = 1.0f / F.d[i];
        int o = 0;
        for (int a = 0; a < 6; ++a) {
            const float ya = F.Y[i][a] * di;
            for (int bb = 0; bb <= a; ++bb) S21[o++] += ya * F.Y[i][bb];
            b6[a] += ya * F.Y[i][6];
        }
    }
}

/* lambda_g = L^-T D^-1 (yr - Yu z);  cforce on own bodies = J^T lambda */
PD_HDN void backsolve_group(const GroupSys& G, const GroupFac& F, const float* z, float* cfA, float* cfB) {
    const int n = G.n;
    float lam[PD_GMAX];
    for (int i = 0; i < n; ++i) {
        float s = F.Y[i][6];
        for (int k = 0; k < 6; ++k) s -= F.Y[i][k] * z[k];
        lam[i] = s / F.d[i];
    }
    for (int i = n - 1; i >= 0; --i) { float s = lam[i]; for (int k = i + 1; k < n; ++k) s -= F.L[tri(k, i)] * lam[k]; lam[i] = s; }
    for (int k = 0; k < 6; ++k) { cfA[k] = 0; cfB[k] = 0; }
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 6; ++k) cfA[k] += G.JA[i][k] * lam[i];
        if (G.hasB) for (int k = 0; k < 6; ++k) cfB[k] += G.JB[i][k] * lam[i];
    }
}

/* util.cpp dxStepBody (finite rotation mode 1, no axis) */
PD_HD void integrate_body(Body& b, float h) {
    b.fr.p.x += h * b.v.x; b.fr.p.y += h * b.v.y; b.fr.p.z += h * b.v.z;
    const float wlen = sqrtf(b.w.x * b.w.x + b.w.y * b.w.y + b.w.z * b.w.z);
    h *= 0.5f;
    const float theta = wlen * h;
    Quat q; q.w = cosf(theta);
    const float sinc = (fabsf(theta) < 1.0e-4f) ? 1.0f - theta * theta * 0.166666666666666666667f : sinf(theta) / theta;
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

/* dWorldStep for the car's island */
PD_HDN void world_step(const PdCarParams& P, CarCtx& X) {
    const float h = X.dt, hinv = 1.0f / h;
    BodyDyn dyn[PD_NUM_BODIES];
    for (int i = 0; i < PD_NUM_BODIES; ++i) {
        Body& b = X.b[i];
        const Sym3 I = rot_diag(b.fr, b.I);
        dyn[i].invI = rot_diag(b.fr, v3(1.0f / b.I.x, 1.0f / b.I.y, 1.0f / b.I.z));
        dyn[i].invMass = 1.0f / b.mass;
        b.T += gyro_torque(b, I, h);
        b.F.y += b.mass * P.gravityY;
        dyn[i].t1[0] = b.F.x * dyn[i].invMass + b.v.x * hinv; dyn[i].t1[1] = b.F.y * dyn[i].invMass + b.v.y * hinv; dyn[i].t1[2] = b.F.z * dyn[i].invMass + b.v.z * hinv;
        const V3 a = sym3_mul(dyn[i].invI, b.T);
        dyn[i].t1[3] = a.x + b.w.x * hinv; dyn[i].t1[4] = a.y + b.w.y * hinv; dyn[i].t1[5] = a.z + b.w.z * hinv;
    }
    float S21[21], b6[6];
    for (int k = 0; k < 21; ++k) S21[k] = 0;
    for (int k = 0; k < 6; ++k) b6[k] = 0;
    GroupSys G[4]; GroupFac F[4];
    const Body& C = X.b[PD_BODY_CHASSIS];
    build_tank(P, X.b[PD_BODY_TANK], C, hinv, G[0]);
    factor_group(G[0], dyn[PD_BODY_TANK], dyn[PD_BODY_TANK], dyn[PD_BODY_CHASSIS], hinv, F[0], S21, b6);
    for (int s = 0; s < 2; ++s) {
        build_strut(P, P.strut[s], C, X.b[PD_BODY_HUB0 + 2 * s], X.b[PD_BODY_STRUT0 + 2 * s], X.steerAnchor1[s], X.steerAnchor2[s], hinv, X.dballErp, X.dballCfm, G[1 + s]);
        factor_group(G[1 + s], dyn[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_STRUT0 + 2 * s], dyn[PD_BODY_CHASSIS], hinv, F[1 + s], S21, b6);
    }
    build_axle(P, C, X.b[PD_BODY_AXLE], hinv, X.dballErp, X.dballCfm, G[3]);
    factor_group(G[3], dyn[PD_BODY_AXLE], dyn[PD_BODY_AXLE], dyn[PD_BODY_CHASSIS], hinv, F[3], S21, b6);
    /* S += M_C (mass on the linear diagonal, world inertia on the angular block) */
    {
        const Sym3 Ic = rot_diag(C.fr, C.I);
        S21[0] += C.mass; S21[2] += C.mass; S21[5] += C.mass;
        S21[9] += Ic.xx; S21[13] += Ic.xy; S21[14] += Ic.yy; S21[18] += Ic.xz; S21[19] += Ic.yz; S21[20] += Ic.zz;
    }
    /* 6x6 LDL^T solve S z = b */
    float z[6];
    {
        float Lm[6][6], d[6];
        for (int i = 0; i < 6; ++i) {
            for (int j = 0; j < i; ++j) { float s = S21[i * (i + 1) / 2 + j]; for (int k = 0; k < j; ++k) s -= Lm[i][k] * Lm[j][k] * d[k]; Lm[i][j] = s / d[j]; }
            float s = S21[i * (i + 1) / 2 + i]; for (int k = 0; k < i; ++k) s -= Lm[i][k] * Lm[i][k] * d[k];
            d[i] = s;
        }
        for (int i = 0; i < 6; ++i) { float s = b6[i]; for (int k = 0; k < i; ++k) s -= Lm[i][k] * z[k]; z[i] = s; }
        for (int i = 0; i < 6; ++i) z[i] /= d[i];
        for (int i = 5; i >= 0; --i) { float s = z[i]; for (int k = i + 1; k < 6; ++k) s -= Lm[k][i] * z[k]; z[i] = s; }
    }
    /* own bodies */
    float cfA[6], cfB[6];
    backsolve_group(G[0], F[0], z, cfA, cfB); apply_update(X.b[PD_BODY_TANK], dyn[PD_BODY_TANK], cfA, h);
    for (int s = 0; s < 2; ++s) {
        backsolve_group(G[1 + s], F[1 + s], z, cfA, cfB);
        apply_update(X.b[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_HUB0 + 2 * s], cfA, h);
        apply_update(X.b[PD_BODY_STRUT0 + 2 * s], dyn[PD_BODY_STRUT0 + 2 * s], cfB, h);
    }
    backsolve_group(G[3], F[3], z, cfA, cfB); apply_update(X.b[PD_BODY_AXLE], dyn[PD_BODY_AXLE], cfA, h);
    /* chassis: z = M_C^-1 U^T lambda  ->  dv = h (M_C^-1 f + z) */
    {
        Body& Cb = X.b[PD_BODY_CHASSIS]; const BodyDyn& d = dyn[PD_BODY_CHASSIS];
        const float imh = h * d.invMass;
        Cb.v.x += Cb.F.x * imh + h * z[0]; Cb.v.y += Cb.F.y * imh + h * z[1]; Cb.v.z += Cb.F.z * imh + h * z[2];
        const V3 dw = sym3_mul(d.invI, v3(Cb.T.x * h, Cb.T.y * h, Cb.T.z * h));
        Cb.w.x += dw.x + h * z[3]; Cb.w.y += dw.y + h * z[4]; Cb.w.z += dw.z + h * z[5];
    }
    for (int i = 0; i < PD_NUM_BODIES; ++i) { integrate_body(X.b[i], h); X.b[i].F = v3(0, 0, 0); X.b[i].T = v3(0, 0, 0); }
}

} // namespace pd
