/*
 * pd_solver2.h -- the constraint solve of dWorldStep as REGISTER-RESIDENT, sparsity-aware, fully unrolled
 * code: no scratch arrays in shared or local memory addressed by loop counters, no padded rows, no stored Jacobian blocks.
 *
 * Same linear system as pd_solver.h (the reference's ODE 0.16.3 step: (J M^-1 J^T + CFM/h) lambda = c/h - J (v/h + M^-1 f),
 * rows restated in oracle/ode_restate/ode_core.h) and the same block-arrow elimination around the chassis
 *     (M_C + sum_g U_g^T A_g^-1 U_g) z = sum_g U_g^T A_g^-1 r_g ,   lambda_g = A_g^-1 (r_g - U_g z),
 * but every group is assembled in CLOSED FORM from the structure of its joints:
 *
 *   strut group (11 rows, own bodies hub H and strut body B), rows ordered [slider angular 3 | slider linear 2 | dball 3 | ball 3]:
 *     slider angular   H: (0, -e_k)        B: (0, +e_k)       chassis: 0
 *     slider linear    H: (-p, c x p / 2)  B: (p, c x p / 2)  chassis: 0
 *     dball l          H: (-q_l, q_l x a2) B: 0               chassis: (q_l, a1 x q_l)
 *     ball k           H: 0                B: (-e_k, e_k x b2) chassis: (e_k, -(e_k x b1))
 *   so A = J M^-1 J^T is written down entry by entry from a handful of 3-vectors (inverse inertia applied to the lever
 *   vectors), the 5 slider rows have no chassis part (their rows of L^-1 U stay zero: ordered first, the forward substitution
 *   of the 6 chassis columns and the Schur accumulation run over 6 rows instead of 11), and the ball / dball blocks do not
 *   couple directly (A[ball][dball] = 0 before fill-in).
 *   single-body groups (tank: fixed joint, 6 rows; rear axle: 5 dball rows): dense 6-vectors per row, one own body.
 *
 * The factorisation is a right-looking L D L^T over compile-time indices (every update of one pivot step is independent:
 * instruction-level parallelism instead of the dependent load-multiply-subtract chain of a row-by-row form), one reciprocal
 * per pivot, explicit fused multiply-adds (PD_FMA: the library is built with -fmad=false for the bit-exact ray / flag code;
 * here contraction only removes roundings).  Elimination order and rounding differ from the oracle's dense row-by-row
 * factorisation; the solution agrees within the conditioning of the system (tests: single-tick parity at 1e-4).
 */
#pragma once
#include "pd_solver.h"
#ifndef PD_SOLVER2
#define PD_SOLVER2 1      /* 0: the scratch-array solver of pd_solver.h (kept for A/B measurements) */
#endif

namespace pd {

#if defined(__CUDA_ARCH__)
#define PD_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define PD_RCP(x) __frcp_rn(x)
#else
#define PD_FMA(a, b, c) fmaf((a), (b), (c))
#define PD_RCP(x) (1.0f / (x))
#endif

PD_HD float fdot(V3 a, V3 b) { return PD_FMA(a.z, b.z, PD_FMA(a.y, b.y, a.x * b.x)); }
PD_HD V3 fcross(V3 a, V3 b) { return v3(PD_FMA(a.y, b.z, -(a.z * b.y)), PD_FMA(a.z, b.x, -(a.x * b.z)), PD_FMA(a.x, b.y, -(a.y * b.x))); }
PD_HD V3 fsym(const Sym3& m, V3 v) {
    return v3(PD_FMA(m.xz, v.z, PD_FMA(m.xy, v.y, m.xx * v.x)), PD_FMA(m.yz, v.z, PD_FMA(m.yy, v.y, m.xy * v.x)), PD_FMA(m.zz, v.z, PD_FMA(m.yz, v.y, m.xz * v.x)));
}
PD_HD float vget(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
PD_HD float symget(const Sym3& m, int i, int j) {
    const int a = i < j ? i : j, b = i < j ? j : i;
    return a == 0 ? (b == 0 ? m.xx : (b == 1 ? m.xy : m.xz)) : (a == 1 ? (b == 1 ? m.yy : m.yz) : m.zz);
}
PD_HD V3 ecross(int k, V3 b) { return k == 0 ? v3(0.0f, -b.z, b.y) : (k == 1 ? v3(b.z, 0.0f, -b.x) : v3(-b.y, b.x, 0.0f)); }   /* e_k x b */
PD_HD V3 eunit(int k) { return v3(k == 0 ? 1.0f : 0.0f, k == 1 ? 1.0f : 0.0f, k == 2 ? 1.0f : 0.0f); }
PD_HD float t1lin_dot(const BodyDyn& d, V3 l) { return PD_FMA(l.z, d.t1[2], PD_FMA(l.y, d.t1[1], l.x * d.t1[0])); }
PD_HD float t1ang_dot(const BodyDyn& d, V3 a) { return PD_FMA(a.z, d.t1[5], PD_FMA(a.y, d.t1[4], a.x * d.t1[3])); }

/* ---------------- dense kernels over compile-time indices (packed lower triangle, row-major: (i,j) at i(i+1)/2 + j) ---------------- */
#define PD_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))

/* right-looking L D L^T in place: on return A(i,j), j < i, holds L_ij and dinv[i] = 1 / D_i (the diagonal slots keep D_i) */
template <int N> PD_HD void ldlt_inplace(float* A, float* dinv) {
    PD_UNROLL
    for (int j = 0; j < N; ++j) {
        const float inv = PD_RCP(A[PD_TRI(j, j)]);
        dinv[j] = inv;
        float l[N];
        PD_UNROLL
        for (int i = j + 1; i < N; ++i) l[i] = A[PD_TRI(i, j)] * inv;
        PD_UNROLL
        for (int i = j + 1; i < N; ++i) {
            PD_UNROLL
            for (int k = j + 1; k <= i; ++k) A[PD_TRI(i, k)] = PD_FMA(-l[i], A[PD_TRI(k, j)], A[PD_TRI(i, k)]);
        }
        PD_UNROLL
        for (int i = j + 1; i < N; ++i) A[PD_TRI(i, j)] = l[i];
    }
}
/* y <- L^-1 y for rows FIRST..N-1 (rows before FIRST are known to be zero and stay zero) */
template <int N, int FIRST> PD_HD void lower_solve(const float* A, float* y) {
    PD_UNROLL
    for (int j = FIRST; j < N; ++j) {
        PD_UNROLL
        for (int i = j + 1; i < N; ++i) y[i] = PD_FMA(-A[PD_TRI(i, j)], y[j], y[i]);
    }
}
/* x <- L^-T x */
template <int N> PD_HD void upper_solve(const float* A, float* x) {
    PD_UNROLL
    for (int j = N - 1; j > 0; --j) {
        PD_UNROLL
        for (int i = 0; i < j; ++i) x[i] = PD_FMA(-A[PD_TRI(j, i)], x[j], x[i]);
    }
}
/* S += sum_i w_i y_i y_i^T (6x6 packed lower), b += sum_i w_i y_i r_i, over NR rows whose chassis part is y[i][0..5] */
template <int NR> PD_HD void schur_accumulate(const float (*Yu)[6], const float* yr, const float* w, float* S21, float* b6) {
    PD_UNROLL
    for (int i = 0; i < NR; ++i) {
        PD_UNROLL
        for (int a = 0; a < 6; ++a) {
            const float ya = Yu[i][a] * w[i];
            PD_UNROLL
            for (int c = 0; c <= a; ++c) S21[PD_TRI(a, c)] = PD_FMA(ya, Yu[i][c], S21[PD_TRI(a, c)]);
            b6[a] = PD_FMA(ya, yr[i], b6[a]);
        }
    }
}

/* What a group keeps between its factorisation and its back-substitution (the chassis unknown z comes in between) lives in ONE
 * flat array of PD_GSYS_WORDS floats per thread, viewed as a strut system or as a single-body system: every index below is a
 * compile-time constant after inlining, so the array is register-allocated element by element (and shared by the two views, which
 * are never live together on one lane). */
#define PD_GSYS_WORDS 157
struct StrutSys {        /* view: A 66 | dinv 11 | Yu 6x6 | yr 11 | p cp q cq 12 | dq 9 | dw 9 | b2 3 */
    float* R;
    PD_HD float& A(int k) const { return R[k]; }
    PD_HD float& dinv(int i) const { return R[66 + i]; }
    PD_HD float& Yu(int i, int c) const { return R[77 + i * 6 + c]; }
    PD_HD float& yr(int i) const { return R[113 + i]; }
    PD_HD V3 vec(int k) const { return v3(R[124 + 3 * k], R[125 + 3 * k], R[126 + 3 * k]); }
    PD_HD void vec(int k, V3 v) const { R[124 + 3 * k] = v.x; R[125 + 3 * k] = v.y; R[126 + 3 * k] = v.z; }
    PD_HD V3 p() const { return vec(0); } PD_HD V3 cp() const { return vec(1); } PD_HD V3 q() const { return vec(2); } PD_HD V3 cq() const { return vec(3); }
    PD_HD V3 dq(int l) const { return vec(4 + l); } PD_HD V3 dw(int l) const { return vec(7 + l); } PD_HD V3 b2() const { return vec(10); }
};
struct SingleSys {       /* view: A 21 | dinv 6 | Yu 6x6 | yr 6 | J 6x6 */
    float* R;
    PD_HD float& A(int k) const { return R[k]; }
    PD_HD float& dinv(int i) const { return R[21 + i]; }
    PD_HD float& Yu(int i, int c) const { return R[27 + i * 6 + c]; }
    PD_HD float& yr(int i) const { return R[63 + i]; }
    PD_HD float& J(int i, int k) const { return R[69 + i * 6 + k]; }
};

/* ---------------- strut group ---------------- */
/* rows + A + r + factorisation + forward substitution; adds this group's share of the chassis Schur system.
 * C chassis, H hub, B strut body; dH / dB / dC their BodyDyn; steerA1 / steerA2: anchors of the steering link (dball 2). */
PD_HD void strut_factor(const PdCarParams& P, const PdStrut& St, const Body& C, const Body& H, const Body& B, V3 steerA1, V3 steerA2,
                        const BodyDyn& dH, const BodyDyn& dB, const BodyDyn& dC, float hinv, float dballErp, float dballCfm, const StrutSys& G, float* S21, float* b6) {
    const float eW = P.worldCFM * hinv, eD = dballCfm * hinv;
    const float mH = dH.invMass, mB = dB.invMass, mHB = mH + mB;
    const Sym3& IH = dH.invI; const Sym3& IB = dB.invI;
    float c[11];
    V3 Ulin[3], Uang[3];      /* chassis part of the dball rows */
    V3 gp_, gcp, gq_, gcq, dq[3], dw[3];
    /* ---- slider (b0 = strut body, b1 = hub): 3 angular rows (0..2), 2 linear rows (3, 4) ---- */
    {
        Quat qrel; qrel.w = St.sliderQrel[0]; qrel.x = St.sliderQrel[1]; qrel.y = St.sliderQrel[2]; qrel.z = St.sliderQrel[3];
        fixed_orientation_c(B, H, qrel, hinv * P.worldERP * 2.0f, c);
        V3 cc = H.fr.p - B.fr.p;
        const V3 ax1 = rot(B.fr, v3(St.sliderAxis1[0], St.sliderAxis1[1], St.sliderAxis1[2]));
        plane_space(ax1, gp_, gq_);
        gcp = cross(cc, gp_) * 0.5f; gcq = cross(cc, gq_) * 0.5f;
        const V3 ofs = rot(H.fr, v3(St.sliderOffset[0], St.sliderOffset[1], St.sliderOffset[2]));
        cc = cc + ofs;
        const float k = hinv * P.worldERP;
        c[3] = k * dot(gp_, cc); c[4] = k * dot(gq_, cc);
    }
    /* ---- dball links (b0 = chassis, b1 = hub): rows 5..7 ---- */
    PD_UNROLL
    for (int l = 0; l < 3; ++l) {
        V3 a1 = v3(St.link[l].anchor1[0], St.link[l].anchor1[1], St.link[l].anchor1[2]);
        V3 a2 = v3(St.link[l].anchor2[0], St.link[l].anchor2[1], St.link[l].anchor2[2]);
        if (l == 2) { a1 = steerA1; a2 = steerA2; }
        float J0[6], J1[6], cl;
        row_dball(C, H, a1, a2, St.link[l].distance, hinv * dballErp, J0, J1, cl);
        Ulin[l] = v3(J0[0], J0[1], J0[2]); Uang[l] = v3(J0[3], J0[4], J0[5]);
        dq[l] = v3(-J1[0], -J1[1], -J1[2]); dw[l] = v3(J1[3], J1[4], J1[5]);
        c[5 + l] = cl;
    }
    /* ---- ball (b0 = chassis, b1 = strut body): rows 8..10 ---- */
    const V3 b1 = rot(C.fr, v3(St.ballAnchor1[0], St.ballAnchor1[1], St.ballAnchor1[2]));
    const V3 b2 = rot(B.fr, v3(St.ballAnchor2[0], St.ballAnchor2[1], St.ballAnchor2[2]));
    {
        const float k = hinv * P.worldERP;
        c[8] = k * (b2.x + B.fr.p.x - b1.x - C.fr.p.x); c[9] = k * (b2.y + B.fr.p.y - b1.y - C.fr.p.y); c[10] = k * (b2.z + B.fr.p.z - b1.z - C.fr.p.z);
    }
    /* ---- A = J_H M_H^-1 J_H^T + J_B M_B^-1 J_B^T + cfm / h, entry by entry ---- */
    float A[66];
    const V3 IHcp = fsym(IH, gcp), IBcp = fsym(IB, gcp), IHcq = fsym(IH, gcq), IBcq = fsym(IB, gcq);
    PD_UNROLL
    for (int i = 0; i < 3; ++i) {
        PD_UNROLL
        for (int j = 0; j <= i; ++j) A[PD_TRI(i, j)] = symget(IH, i, j) + symget(IB, i, j) + (i == j ? eW : 0.0f);
    }
    {
        const V3 gp = IBcp - IHcp, gq = IBcq - IHcq;
        A[PD_TRI(3, 0)] = gp.x; A[PD_TRI(3, 1)] = gp.y; A[PD_TRI(3, 2)] = gp.z;
        A[PD_TRI(4, 0)] = gq.x; A[PD_TRI(4, 1)] = gq.y; A[PD_TRI(4, 2)] = gq.z;
        const V3 hp = IHcp + IBcp, hq = IHcq + IBcq;
        A[PD_TRI(3, 3)] = PD_FMA(mHB, fdot(gp_, gp_), fdot(gcp, hp)) + eW;
        A[PD_TRI(4, 3)] = PD_FMA(mHB, fdot(gq_, gp_), fdot(gcq, hp));
        A[PD_TRI(4, 4)] = PD_FMA(mHB, fdot(gq_, gq_), fdot(gcq, hq)) + eW;
    }
    V3 Iw[3];
    PD_UNROLL
    for (int l = 0; l < 3; ++l) {
        Iw[l] = fsym(IH, dw[l]);
        A[PD_TRI(5 + l, 0)] = -Iw[l].x; A[PD_TRI(5 + l, 1)] = -Iw[l].y; A[PD_TRI(5 + l, 2)] = -Iw[l].z;
        A[PD_TRI(5 + l, 3)] = PD_FMA(mH, fdot(dq[l], gp_), fdot(dw[l], IHcp));
        A[PD_TRI(5 + l, 4)] = PD_FMA(mH, fdot(dq[l], gq_), fdot(dw[l], IHcq));
        PD_UNROLL
        for (int m = 0; m <= l; ++m) A[PD_TRI(5 + l, 5 + m)] = PD_FMA(mH, fdot(dq[l], dq[m]), fdot(dw[l], Iw[m])) + (l == m ? eD : 0.0f);
    }
    V3 bn[3], In[3];
    PD_UNROLL
    for (int k = 0; k < 3; ++k) {
        bn[k] = ecross(k, b2);
        In[k] = fsym(IB, bn[k]);
        A[PD_TRI(8 + k, 0)] = In[k].x; A[PD_TRI(8 + k, 1)] = In[k].y; A[PD_TRI(8 + k, 2)] = In[k].z;
        A[PD_TRI(8 + k, 3)] = PD_FMA(-mB, vget(gp_, k), fdot(bn[k], IBcp));
        A[PD_TRI(8 + k, 4)] = PD_FMA(-mB, vget(gq_, k), fdot(bn[k], IBcq));
        A[PD_TRI(8 + k, 5)] = 0.0f; A[PD_TRI(8 + k, 6)] = 0.0f; A[PD_TRI(8 + k, 7)] = 0.0f;
        PD_UNROLL
        for (int j = 0; j <= k; ++j) A[PD_TRI(8 + k, 8 + j)] = fdot(bn[k], In[j]) + (k == j ? mB + eW : 0.0f);
    }
    /* ---- right-hand side r_i = c_i / h - J_i (v / h + M^-1 f) ---- */
    float r[11];
    r[0] = PD_FMA(c[0], hinv, dH.t1[3] - dB.t1[3]); r[1] = PD_FMA(c[1], hinv, dH.t1[4] - dB.t1[4]); r[2] = PD_FMA(c[2], hinv, dH.t1[5] - dB.t1[5]);
    r[3] = PD_FMA(c[3], hinv, t1lin_dot(dH, gp_) - t1lin_dot(dB, gp_) - t1ang_dot(dH, gcp) - t1ang_dot(dB, gcp));
    r[4] = PD_FMA(c[4], hinv, t1lin_dot(dH, gq_) - t1lin_dot(dB, gq_) - t1ang_dot(dH, gcq) - t1ang_dot(dB, gcq));
    PD_UNROLL
    for (int l = 0; l < 3; ++l)
        r[5 + l] = PD_FMA(c[5 + l], hinv, t1lin_dot(dH, dq[l]) - t1ang_dot(dH, dw[l]) - t1lin_dot(dC, Ulin[l]) - t1ang_dot(dC, Uang[l]));
    PD_UNROLL
    for (int k = 0; k < 3; ++k)
        r[8 + k] = PD_FMA(c[8 + k], hinv, dB.t1[k] - t1ang_dot(dB, bn[k]) - dC.t1[k] + t1ang_dot(dC, ecross(k, b1)));
    /* ---- factor, forward substitution ---- */
    float dinv[11];
    ldlt_inplace<11>(A, dinv);
    lower_solve<11, 0>(A, r);
    /* chassis columns: rows 5..7 = (q_l, a1 x q_l), rows 8..10 = (e_k, -(e_k x b1)); forward substitution inside rows 5..10 */
    float Yu[6][6];
    PD_UNROLL
    for (int col = 0; col < 6; ++col) {
        float y[11];
        PD_UNROLL
        for (int i = 0; i < 5; ++i) y[i] = 0.0f;
        PD_UNROLL
        for (int l = 0; l < 3; ++l) y[5 + l] = col < 3 ? vget(Ulin[l], col) : vget(Uang[l], col - 3);
        PD_UNROLL
        for (int k = 0; k < 3; ++k) y[8 + k] = col < 3 ? (col == k ? 1.0f : 0.0f) : -vget(ecross(k, b1), col - 3);
        lower_solve<11, 5>(A, y);
        PD_UNROLL
        for (int i = 0; i < 6; ++i) Yu[i][col] = y[5 + i];
    }
    schur_accumulate<6>(Yu, r + 5, dinv + 5, S21, b6);
    /* keep what the back-substitution needs */
    PD_UNROLL
    for (int k = 0; k < 66; ++k) G.A(k) = A[k];
    PD_UNROLL
    for (int i = 0; i < 11; ++i) { G.dinv(i) = dinv[i]; G.yr(i) = r[i]; }
    PD_UNROLL
    for (int i = 0; i < 6; ++i) { PD_UNROLL for (int k = 0; k < 6; ++k) G.Yu(i, k) = Yu[i][k]; }
    G.vec(0, gp_); G.vec(1, gcp); G.vec(2, gq_); G.vec(3, gcq);
    PD_UNROLL
    for (int l = 0; l < 3; ++l) { G.vec(4 + l, dq[l]); G.vec(7 + l, dw[l]); }
    G.vec(10, b2);
}

/* lambda = L^-T D^-1 (yr - Yu z); constraint force / torque on the hub (cfH) and the strut body (cfB) */
PD_HD void strut_backsolve(const StrutSys& G, const float* z, float* cfH, float* cfB) {
    float lam[11];
    PD_UNROLL
    for (int i = 0; i < 5; ++i) lam[i] = G.yr(i) * G.dinv(i);
    PD_UNROLL
    for (int i = 0; i < 6; ++i) {
        float s = G.yr(5 + i);
        PD_UNROLL
        for (int k = 0; k < 6; ++k) s = PD_FMA(-G.Yu(i, k), z[k], s);
        lam[5 + i] = s * G.dinv(5 + i);
    }
    upper_solve<11>(G.R, lam);
    /* hub: sum_l lam_dl (-dq_l, dw_l) + lam_3 (-p, cp) + lam_4 (-q, cq) + (0, -lam_012) */
    V3 fl = G.p() * (-lam[3]) + G.q() * (-lam[4]);
    V3 fa = G.cp() * lam[3] + G.cq() * lam[4] + v3(-lam[0], -lam[1], -lam[2]);
    PD_UNROLL
    for (int l = 0; l < 3; ++l) { fl = fl + G.dq(l) * (-lam[5 + l]); fa = fa + G.dw(l) * lam[5 + l]; }
    cfH[0] = fl.x; cfH[1] = fl.y; cfH[2] = fl.z; cfH[3] = fa.x; cfH[4] = fa.y; cfH[5] = fa.z;
    /* strut body: lam_3 (p, cp) + lam_4 (q, cq) + (0, lam_012) + (-lam_b, lam_b x b2) */
    const V3 lb = v3(lam[8], lam[9], lam[10]);
    const V3 gl = G.p() * lam[3] + G.q() * lam[4] - lb;
    const V3 ga = G.cp() * lam[3] + G.cq() * lam[4] + v3(lam[0], lam[1], lam[2]) + cross(lb, G.b2());
    cfB[0] = gl.x; cfB[1] = gl.y; cfB[2] = gl.z; cfB[3] = ga.x; cfB[4] = ga.y; cfB[5] = ga.z;
}

/* ---------------- groups with ONE own body and up to 6 dense rows (tank: fixed joint; rear axle: dball links) ---------------- */
PD_HD void single_row(const SingleSys& G, int i, V3 jl, V3 ja, V3 ul, V3 ua, float c) {
    G.J(i, 0) = jl.x; G.J(i, 1) = jl.y; G.J(i, 2) = jl.z; G.J(i, 3) = ja.x; G.J(i, 4) = ja.y; G.J(i, 5) = ja.z;
    G.Yu(i, 0) = ul.x; G.Yu(i, 1) = ul.y; G.Yu(i, 2) = ul.z; G.Yu(i, 3) = ua.x; G.Yu(i, 4) = ua.y; G.Yu(i, 5) = ua.z;
    G.yr(i) = c;
}
/* fixed joint tank (b0, own) <-> chassis (b1) (fixed.cpp getInfo2: rows 0-2 linear, 3-5 angular) */
PD_HD void single_rows_tank(const PdCarParams& P, const Body& T, const Body& C, float hinv, const SingleSys& G, float* cfm) {
    const V3 ofs = rot(T.fr, v3(P.tankOffset[0], P.tankOffset[1], P.tankOffset[2]));
    const float k = hinv * P.worldERP;
    const V3 cl = v3(k * (C.fr.p.x - T.fr.p.x + ofs.x), k * (C.fr.p.y - T.fr.p.y + ofs.y), k * (C.fr.p.z - T.fr.p.z + ofs.z));
    Quat qrel; qrel.w = P.tankQrel[0]; qrel.x = P.tankQrel[1]; qrel.y = P.tankQrel[2]; qrel.z = P.tankQrel[3];
    float c3[3]; fixed_orientation_c(T, C, qrel, hinv * P.worldERP * 2.0f, c3);
    PD_UNROLL
    for (int k3 = 0; k3 < 3; ++k3) {
        single_row(G, k3, eunit(k3), ecross(k3, ofs), neg(eunit(k3)), v3(0, 0, 0), vget(cl, k3));
        single_row(G, 3 + k3, v3(0, 0, 0), eunit(k3), v3(0, 0, 0), neg(eunit(k3)), c3[k3]);
    }
    PD_UNROLL
    for (int i = 0; i < 6; ++i) cfm[i] = P.worldCFM;
}
/* dball links chassis (b0) <-> own body (b1: rigid axle, or the hub of a double-wishbone corner); rows beyond n are identity
 * padding.  steer != null: link 4's anchors are the re-seated steering rod (SuspensionDW::setSteerLengthOffset) */
PD_HD void single_rows_links(const PdDBall* links, int n, const Body& C, const Body& Ax, float hinv, float dballErp, float dballCfm, const SingleSys& G, float* cfm,
                             const V3* steer = nullptr) {
    PD_UNROLL
    for (int l = 0; l < 6; ++l) {
        if (l < PD_AXLE_LINKS && l < n) {
            const PdDBall& K = links[l];
            float J0[6], J1[6], c;
            V3 a1 = v3(K.anchor1[0], K.anchor1[1], K.anchor1[2]), a2 = v3(K.anchor2[0], K.anchor2[1], K.anchor2[2]);
            if (steer && l == 4) { a1 = steer[0]; a2 = steer[1]; }
            row_dball(C, Ax, a1, a2, K.distance, hinv * dballErp, J0, J1, c);
            single_row(G, l, v3(J1[0], J1[1], J1[2]), v3(J1[3], J1[4], J1[5]), v3(J0[0], J0[1], J0[2]), v3(J0[3], J0[4], J0[5]), c);
            cfm[l] = dballCfm;
        } else {
            single_row(G, l, v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0), v3(0, 0, 0), 0.0f);
            cfm[l] = 1.0f / hinv;       /* pad row: diagonal cfm / h = 1, nothing else */
        }
    }
}
PD_HD void single_rows_axle(const PdCarParams& P, const Body& C, const Body& Ax, float hinv, float dballErp, float dballCfm, const SingleSys& G, float* cfm) {
    single_rows_links(P.axle.link, P.axle.nLinks, C, Ax, hinv, dballErp, dballCfm, G, cfm);
}
PD_HD void single_factor(const SingleSys& G, const float* cfm, const BodyDyn& dA, const BodyDyn& dC, float hinv, float* S21, float* b6) {
    float MJ[6][6], A[21], r[6], dinv[6], Yu[6][6];
    PD_UNROLL
    for (int i = 0; i < 6; ++i) {
        const V3 a = fsym(dA.invI, v3(G.J(i, 3), G.J(i, 4), G.J(i, 5)));
        MJ[i][0] = G.J(i, 0) * dA.invMass; MJ[i][1] = G.J(i, 1) * dA.invMass; MJ[i][2] = G.J(i, 2) * dA.invMass; MJ[i][3] = a.x; MJ[i][4] = a.y; MJ[i][5] = a.z;
    }
    PD_UNROLL
    for (int i = 0; i < 6; ++i) {
        PD_UNROLL
        for (int j = 0; j <= i; ++j) {
            float s = G.J(i, 0) * MJ[j][0];
            PD_UNROLL
            for (int k = 1; k < 6; ++k) s = PD_FMA(G.J(i, k), MJ[j][k], s);
            A[PD_TRI(i, j)] = s + (i == j ? cfm[i] * hinv : 0.0f);
        }
        float s = G.J(i, 0) * dA.t1[0];
        PD_UNROLL
        for (int k = 1; k < 6; ++k) s = PD_FMA(G.J(i, k), dA.t1[k], s);
        PD_UNROLL
        for (int k = 0; k < 6; ++k) s = PD_FMA(G.Yu(i, k), dC.t1[k], s);
        r[i] = PD_FMA(G.yr(i), hinv, -s);
    }
    ldlt_inplace<6>(A, dinv);
    lower_solve<6, 0>(A, r);
    PD_UNROLL
    for (int col = 0; col < 6; ++col) {
        float y[6];
        PD_UNROLL
        for (int i = 0; i < 6; ++i) y[i] = G.Yu(i, col);
        lower_solve<6, 0>(A, y);
        PD_UNROLL
        for (int i = 0; i < 6; ++i) Yu[i][col] = y[i];
    }
    schur_accumulate<6>(Yu, r, dinv, S21, b6);
    PD_UNROLL
    for (int k = 0; k < 21; ++k) G.A(k) = A[k];
    PD_UNROLL
    for (int i = 0; i < 6; ++i) { G.dinv(i) = dinv[i]; G.yr(i) = r[i]; PD_UNROLL for (int k = 0; k < 6; ++k) G.Yu(i, k) = Yu[i][k]; }
}
PD_HD void single_backsolve(const SingleSys& G, const float* z, float* cfA) {
    float lam[6];
    PD_UNROLL
    for (int i = 0; i < 6; ++i) {
        float s = G.yr(i);
        PD_UNROLL
        for (int k = 0; k < 6; ++k) s = PD_FMA(-G.Yu(i, k), z[k], s);
        lam[i] = s * G.dinv(i);
    }
    upper_solve<6>(G.R, lam);
    PD_UNROLL
    for (int k = 0; k < 6; ++k) {
        float s = G.J(0, k) * lam[0];
        PD_UNROLL
        for (int i = 1; i < 6; ++i) s = PD_FMA(G.J(i, k), lam[i], s);
        cfA[k] = s;
    }
}

/* dWorldStep for the car's island, one thread doing the four groups one after the other (thread-per-car kernel, host build) */
PD_HDN void contacts_solve(const PdCarParams& P, const float* __restrict__ cont, const Body& C, const BodyDyn& dC, const float* __restrict__ S21, const float* __restrict__ b6, float h, bool fresh, float& lifeLeft, float* __restrict__ dmg, float* __restrict__ z);   /* pd_contacts.h */
#define PD_TOPO_FRONT_DW(T) (((T) & 2) != 0)
#define PD_TOPO_REAR_DW(T) (((T) & 1) != 0)
/* body slots a topology owns (include/pd_state.h): chassis, tank, the two front hubs and slot 6 always; the strut bodies with a
 * strut front axle; slot 7 with a double-wishbone rear axle */
template <int TOPO> PD_HD constexpr bool topo_has_body(int i) { return (i == PD_BODY_STRUT0 || i == PD_BODY_STRUT1) ? !PD_TOPO_FRONT_DW(TOPO) : (i == PD_BODY_HUB3 ? PD_TOPO_REAR_DW(TOPO) : true); }
PD_HD bool topo_has_body_rt(int topo, int i) { return (i == PD_BODY_STRUT0 || i == PD_BODY_STRUT1) ? !PD_TOPO_FRONT_DW(topo) : (i == PD_BODY_HUB3 ? PD_TOPO_REAR_DW(topo) : true); }
/* a double-wishbone corner as a single-body group: hub w held by its five links */
PD_HD void dw_factor(const PdCarParams& P, const PdDW& D, const Body& C, const Body& H, const V3* steer, const BodyDyn& dH, const BodyDyn& dC, float hinv, float dballErp, float dballCfm, float* R, float* S21, float* b6) {
    SingleSys G; G.R = R; float cfm[6];
    if (D.multilink) { dballErp = P.worldERP; dballCfm = P.worldCFM; }      /* SuspensionML::setERPCFM is empty: its joints keep the world's values */
    single_rows_links(D.link, PD_DW_LINKS, C, H, hinv, dballErp, dballCfm, G, cfm, steer);
    single_factor(G, cfm, dH, dC, hinv, S21, b6);
}
template <int TOPO = 0>
PD_HDN void world_step2(const PdCarParams& P, Body* b, const V3* steerAnchor1, const V3* steerAnchor2, float dballErp, float dballCfm, float h, const float* cont, bool freshContacts, float& lifeLeft, float* dmg) {
    constexpr bool FDW = PD_TOPO_FRONT_DW(TOPO), RDW = PD_TOPO_REAR_DW(TOPO);
    const float hinv = 1.0f / h;
    BodyDyn dyn[PD_NUM_BODIES];
    PD_UNROLL
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body<TOPO>(i)) body_dyn(b[i], P.gravityY, h, dyn[i]);
    float S21[21], b6[6];
    for (int k = 0; k < 21; ++k) S21[k] = 0;
    for (int k = 0; k < 6; ++k) b6[k] = 0;
    const Body& C = b[PD_BODY_CHASSIS];
    float RS[2][FDW ? 105 : PD_GSYS_WORDS], RT[105], RA[RDW ? 2 : 1][105];
    float cfm[6];
    { SingleSys GT; GT.R = RT; single_rows_tank(P, b[PD_BODY_TANK], C, hinv, GT, cfm); single_factor(GT, cfm, dyn[PD_BODY_TANK], dyn[PD_BODY_CHASSIS], hinv, S21, b6); }
    PD_NOUNROLL
    for (int s = 0; s < 2; ++s) {
        if constexpr (FDW) {
            const V3 st[2] = {steerAnchor1[s], steerAnchor2[s]};
            dw_factor(P, P.dw[s], C, b[PD_BODY_HUB0 + 2 * s], st, dyn[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_CHASSIS], hinv, dballErp, dballCfm, RS[s], S21, b6);
        } else {
            StrutSys GS; GS.R = RS[s];
            strut_factor(P, P.strut[s], C, b[PD_BODY_HUB0 + 2 * s], b[PD_BODY_STRUT0 + 2 * s], steerAnchor1[s], steerAnchor2[s], dyn[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_STRUT0 + 2 * s],
                         dyn[PD_BODY_CHASSIS], hinv, dballErp, dballCfm, GS, S21, b6);
        }
    }
    if constexpr (RDW) {
        PD_NOUNROLL
        for (int s = 0; s < 2; ++s) dw_factor(P, P.dw[2 + s], C, b[PD_BODY_HUB2 + s], nullptr, dyn[PD_BODY_HUB2 + s], dyn[PD_BODY_CHASSIS], hinv, dballErp, dballCfm, RA[s], S21, b6);
    } else { SingleSys GA; GA.R = RA[0]; single_rows_axle(P, C, b[PD_BODY_AXLE], hinv, dballErp, dballCfm, GA, cfm); single_factor(GA, cfm, dyn[PD_BODY_AXLE], dyn[PD_BODY_CHASSIS], hinv, S21, b6); }
    schur_add_chassis(S21, C);
    float z[6];
    if (cont && reinterpret_cast<const int*>(cont)[0] > 0) contacts_solve(P, cont, C, dyn[PD_BODY_CHASSIS], S21, b6, h, freshContacts, lifeLeft, dmg, z);   /* live contact joints: rows on the chassis */
    else solve6(S21, b6, z);
    float cfA[6], cfB[6];
    { SingleSys GT; GT.R = RT; single_backsolve(GT, z, cfA); apply_update(b[PD_BODY_TANK], dyn[PD_BODY_TANK], cfA, h); }
    PD_NOUNROLL
    for (int s = 0; s < 2; ++s) {
        if constexpr (FDW) {
            SingleSys GH; GH.R = RS[s]; single_backsolve(GH, z, cfA); apply_update(b[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_HUB0 + 2 * s], cfA, h);
        } else {
            StrutSys GS; GS.R = RS[s];
            strut_backsolve(GS, z, cfA, cfB);
            apply_update(b[PD_BODY_HUB0 + 2 * s], dyn[PD_BODY_HUB0 + 2 * s], cfA, h);
            apply_update(b[PD_BODY_STRUT0 + 2 * s], dyn[PD_BODY_STRUT0 + 2 * s], cfB, h);
        }
    }
    if constexpr (RDW) {
        PD_NOUNROLL
        for (int s = 0; s < 2; ++s) { SingleSys GH; GH.R = RA[s]; single_backsolve(GH, z, cfA); apply_update(b[PD_BODY_HUB2 + s], dyn[PD_BODY_HUB2 + s], cfA, h); }
    } else { SingleSys GA; GA.R = RA[0]; single_backsolve(GA, z, cfA); apply_update(b[PD_BODY_AXLE], dyn[PD_BODY_AXLE], cfA, h); }
    chassis_update(b[PD_BODY_CHASSIS], dyn[PD_BODY_CHASSIS], z, h);
    PD_UNROLL
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body<TOPO>(i)) { integrate_body(b[i], h); b[i].F = v3(0, 0, 0); b[i].T = v3(0, 0, 0); }
}

} // namespace pd
