/*
 * oracle/ode_restate/ode_core.h -- TEST INFRASTRUCTURE (oracle only; never linked into the product).
 *
 * CPU restatement of the part of the hot path that lives in a third-party dependency ABSENT from
 * /root/reference: Open Dynamics Engine 0.16.3, single precision (pinned by
 * thirdparty/ode/include/ode/version.h:4 and thirdparty/ode/ode_flags.txt:1-5; the import library is
 * listed in .MISSING_LARGE_BLOBS:10).  PARITY UNPINNED: the reference ships no test, golden vector or
 * binary of ODE, so this file restates ODE's published algorithm (ode/src/step.cpp dxStepIsland,
 * ode/src/joints/{fixed,ball,slider,dball}.cpp getInfo2, ode/src/util.cpp dxStepBody,
 * ode/src/rotation.cpp, ode/src/mass.cpp, ode/src/fastldlt.c / lcp.cpp unbounded short-circuit) and is
 * anchored on the reference's own call sites:
 *   Physics/ODE/PhysicsEngineODE.cpp:14-29 (world params), :216-224 (dWorldStep),
 *   Physics/ODE/RigidBodyODE.cpp:11-18,64-70,101-127,184-270,
 *   Physics/ODE/JointODE.cpp:21-88.
 *
 * Everything is `float` exactly as ODE built with dSINGLE.
 */
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace oder {

typedef float dReal;

struct Body {
    dReal pos[3] = {0, 0, 0};
    dReal q[4] = {1, 0, 0, 0};
    dReal R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   /* row-major 3x3, world = R * local */
    dReal lvel[3] = {0, 0, 0}, avel[3] = {0, 0, 0};
    dReal facc[3] = {0, 0, 0}, tacc[3] = {0, 0, 0};
    dReal mass = 1, I[3] = {1, 1, 1};            /* diagonal body-frame inertia (dMassSetBoxTotal) */
    int index = -1;                              /* position in World::bodies */
};

enum JointType { J_FIXED = 0, J_BALL = 1, J_SLIDER = 2, J_DBALL = 3 };

struct Joint {
    JointType type;
    Body* b0 = nullptr;
    Body* b1 = nullptr;
    dReal anchor1[3] = {0, 0, 0}, anchor2[3] = {0, 0, 0}; /* body-local anchors (ball, dball) */
    dReal erp = 0, cfm = 0;                               /* ball / dball / fixed own copies */
    dReal offset[3] = {0, 0, 0};                          /* fixed: in b0 frame; slider: in b1 frame */
    dReal qrel[4] = {1, 0, 0, 0};                         /* fixed, slider */
    dReal axis1[3] = {1, 0, 0};                           /* slider axis in b0 frame */
    dReal targetDistance = 0;                             /* dball */
    int rows() const { return type == J_FIXED ? 6 : type == J_BALL ? 3 : type == J_SLIDER ? 5 : 1; }
};

/* dxJointContact between a body and the static world (dJointAttach(j, body, 0)): geometry fixed at creation
 * (ode/src/joints/contact.cpp getInfo2); surface parameters of PhysicsEngineODE::onCollision (PhysicsEngineODE.cpp:299-322) */
struct ContactJoint {
    Body* b0 = nullptr;
    dReal pos[3], normal[3], depth;
    dReal mu, bounce, soft_cfm, soft_erp;      /* soft_erp < 0: world ERP (mode without dContactSoftERP) */
};

struct World {
    dReal gravity[3] = {0, 0, 0};
    dReal erp = 0.2f, cfm = 1e-5f;
    dReal contactMaxCorrectingVel = 3.0f, contactSurfaceLayer = 0.0f;     /* PhysicsEngineODE.cpp:26-27 */
    std::vector<Body*> bodies;
    std::vector<Joint*> joints;
    std::vector<ContactJoint> contacts;          /* the joints of both contact groups that are alive in this step */
    int pgsIterations = 50;
    bool keepSystem = false;                     /* tests: keep A (before factoring) and rhs of the last step's bilateral block */
    std::vector<dReal> lastA, lastRhs;
    /* diagnostics of the last step */
    int last_m = 0;
    std::vector<dReal> last_lambda;
};

/* ---------------- small vector helpers (ODE odemath.h naming) ---------------- */
inline dReal dot3(const dReal* a, const dReal* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline void cross3(dReal* r, const dReal* a, const dReal* b) {
    r[0] = a[1] * b[2] - a[2] * b[1]; r[1] = a[2] * b[0] - a[0] * b[2]; r[2] = a[0] * b[1] - a[1] * b[0];
}
/* dMultiply0_331: r = R * v */
inline void mul0_331(dReal* r, const dReal* R, const dReal* v) {
    r[0] = dot3(R + 0, v); r[1] = dot3(R + 3, v); r[2] = dot3(R + 6, v);
}
/* dMultiply1_331: r = R^T * v */
inline void mul1_331(dReal* r, const dReal* R, const dReal* v) {
    r[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    r[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    r[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}
inline void normalize3(dReal* a) {
    dReal l = dot3(a, a);
    if (l > 0) { l = 1.0f / sqrtf(l); a[0] *= l; a[1] *= l; a[2] *= l; } else { a[0] = 1; a[1] = 0; a[2] = 0; }
}
inline void normalize4(dReal* a) {
    dReal l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3];
    if (l > 0) { l = 1.0f / sqrtf(l); a[0] *= l; a[1] *= l; a[2] *= l; a[3] *= l; } else { a[0] = 1; a[1] = a[2] = a[3] = 0; }
}
/* rotation.cpp dQtoR */
inline void q_to_R(const dReal* q, dReal* R) {
    dReal qq1 = 2 * q[1] * q[1], qq2 = 2 * q[2] * q[2], qq3 = 2 * q[3] * q[3];
    R[0] = 1 - qq2 - qq3;                 R[1] = 2 * (q[1] * q[2] - q[0] * q[3]); R[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
    R[3] = 2 * (q[1] * q[2] + q[0] * q[3]); R[4] = 1 - qq1 - qq3;                 R[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
    R[6] = 2 * (q[1] * q[3] - q[0] * q[2]); R[7] = 2 * (q[2] * q[3] + q[0] * q[1]); R[8] = 1 - qq1 - qq2;
}
/* rotation.cpp dRtoQ */
inline void R_to_q(const dReal* R, dReal* q) {
#define _R(i, j) R[(i) * 3 + (j)]
    dReal tr = _R(0, 0) + _R(1, 1) + _R(2, 2), s;
    if (tr >= 0) {
        s = sqrtf(tr + 1); q[0] = 0.5f * s; s = 0.5f * (1.0f / s);
        q[1] = (_R(2, 1) - _R(1, 2)) * s; q[2] = (_R(0, 2) - _R(2, 0)) * s; q[3] = (_R(1, 0) - _R(0, 1)) * s;
    } else {
        int c = 0;
        if (_R(1, 1) > _R(0, 0)) { c = (_R(2, 2) > _R(1, 1)) ? 2 : 1; } else if (_R(2, 2) > _R(0, 0)) c = 2;
        if (c == 0) {
            s = sqrtf((_R(0, 0) - (_R(1, 1) + _R(2, 2))) + 1); q[1] = 0.5f * s; s = 0.5f * (1.0f / s);
            q[2] = (_R(0, 1) + _R(1, 0)) * s; q[3] = (_R(2, 0) + _R(0, 2)) * s; q[0] = (_R(2, 1) - _R(1, 2)) * s;
        } else if (c == 1) {
            s = sqrtf((_R(1, 1) - (_R(2, 2) + _R(0, 0))) + 1); q[2] = 0.5f * s; s = 0.5f * (1.0f / s);
            q[3] = (_R(1, 2) + _R(2, 1)) * s; q[1] = (_R(0, 1) + _R(1, 0)) * s; q[0] = (_R(0, 2) - _R(2, 0)) * s;
        } else {
            s = sqrtf((_R(2, 2) - (_R(0, 0) + _R(1, 1))) + 1); q[3] = 0.5f * s; s = 0.5f * (1.0f / s);
            q[1] = (_R(2, 0) + _R(0, 2)) * s; q[2] = (_R(1, 2) + _R(2, 1)) * s; q[0] = (_R(1, 0) - _R(0, 1)) * s;
        }
    }
#undef _R
}
/* quaternion products (rotation.cpp): qmul0 = b*c, qmul1 = inv(b)*c, qmul2 = b*inv(c) */
inline void qmul0(dReal* a, const dReal* b, const dReal* c) {
    a[0] = b[0] * c[0] - b[1] * c[1] - b[2] * c[2] - b[3] * c[3];
    a[1] = b[0] * c[1] + b[1] * c[0] + b[2] * c[3] - b[3] * c[2];
    a[2] = b[0] * c[2] + b[2] * c[0] + b[3] * c[1] - b[1] * c[3];
    a[3] = b[0] * c[3] + b[3] * c[0] + b[1] * c[2] - b[2] * c[1];
}
inline void qmul1(dReal* a, const dReal* b, const dReal* c) {
    a[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    a[1] = b[0] * c[1] - b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    a[2] = b[0] * c[2] - b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    a[3] = b[0] * c[3] - b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
}
inline void qmul2(dReal* a, const dReal* b, const dReal* c) {
    a[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    a[1] = -b[0] * c[1] + b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    a[2] = -b[0] * c[2] + b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    a[3] = -b[0] * c[3] + b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
}
/* odemath.cpp dPlaneSpace */
inline void plane_space(const dReal* n, dReal* p, dReal* q) {
    if (fabsf(n[2]) > 0.70710678118654752440f) {
        dReal a = n[1] * n[1] + n[2] * n[2], k = 1.0f / sqrtf(a);
        p[0] = 0; p[1] = -n[2] * k; p[2] = n[1] * k;
        q[0] = a * k; q[1] = -n[0] * p[2]; q[2] = n[0] * p[1];
    } else {
        dReal a = n[0] * n[0] + n[1] * n[1], k = 1.0f / sqrtf(a);
        p[0] = -n[1] * k; p[1] = n[0] * k; p[2] = 0;
        q[0] = -n[2] * p[1]; q[1] = n[2] * p[0]; q[2] = a * k;
    }
}

/* ---------------- body API used through Physics/ODE/RigidBodyODE.cpp ---------------- */
inline void body_set_mass_box(Body& b, dReal m, dReal lx, dReal ly, dReal lz) { /* mass.cpp dMassSetBoxTotal */
    b.mass = m;
    b.I[0] = m / 12.0f * (ly * ly + lz * lz);
    b.I[1] = m / 12.0f * (lx * lx + lz * lz);
    b.I[2] = m / 12.0f * (lx * lx + ly * ly);
}
/* rotation.cpp dxOrthogonalizeR: Gram-Schmidt on the ROWS of R, third row = row0 x row1 */
inline void orthogonalize_R(dReal* m) {
    dReal n0 = dot3(m, m);
    if (n0 != 1.0f) normalize3(m);
    dReal proj = dot3(m, m + 3);
    if (proj != 0) { m[3] -= proj * m[0]; m[4] -= proj * m[1]; m[5] -= proj * m[2]; }
    dReal n1 = dot3(m + 3, m + 3);
    if (n1 != 1.0f) normalize3(m + 3);
    cross3(m + 6, m, m + 3);
}
inline void body_set_rotation(Body& b, const dReal* R) { /* ode.cpp dBodySetRotation: R orthogonalised, q from the raw input */
    memcpy(b.R, R, sizeof(dReal) * 9);
    orthogonalize_R(b.R);
    R_to_q(R, b.q); normalize4(b.q);
}
inline void body_rel_point_pos(const Body& b, const dReal* p, dReal* r) { /* dBodyGetRelPointPos */
    dReal t[3]; mul0_331(t, b.R, p); r[0] = t[0] + b.pos[0]; r[1] = t[1] + b.pos[1]; r[2] = t[2] + b.pos[2];
}
inline void body_pos_rel_point(const Body& b, const dReal* p, dReal* r) { /* dBodyGetPosRelPoint */
    dReal t[3] = {p[0] - b.pos[0], p[1] - b.pos[1], p[2] - b.pos[2]}; mul1_331(r, b.R, t);
}
inline void body_point_vel(const Body& b, const dReal* p, dReal* r) { /* dBodyGetPointVel */
    dReal t[3] = {p[0] - b.pos[0], p[1] - b.pos[1], p[2] - b.pos[2]};
    dReal c[3]; cross3(c, b.avel, t);
    r[0] = b.lvel[0] + c[0]; r[1] = b.lvel[1] + c[1]; r[2] = b.lvel[2] + c[2];
}
inline void body_rel_point_vel(const Body& b, const dReal* prel, dReal* r) { /* dBodyGetRelPointVel */
    dReal t[3]; mul0_331(t, b.R, prel);
    dReal c[3]; cross3(c, b.avel, t);
    r[0] = b.lvel[0] + c[0]; r[1] = b.lvel[1] + c[1]; r[2] = b.lvel[2] + c[2];
}
inline void body_add_force_at_pos(Body& b, const dReal* f, const dReal* p) { /* dBodyAddForceAtPos */
    b.facc[0] += f[0]; b.facc[1] += f[1]; b.facc[2] += f[2];
    dReal t[3] = {p[0] - b.pos[0], p[1] - b.pos[1], p[2] - b.pos[2]};
    dReal c[3]; cross3(c, t, f);
    b.tacc[0] += c[0]; b.tacc[1] += c[1]; b.tacc[2] += c[2];
}
inline void body_add_force_at_rel_pos(Body& b, const dReal* f, const dReal* prel) { /* dBodyAddForceAtRelPos */
    dReal t[3]; mul0_331(t, b.R, prel);
    b.facc[0] += f[0]; b.facc[1] += f[1]; b.facc[2] += f[2];
    dReal c[3]; cross3(c, t, f);
    b.tacc[0] += c[0]; b.tacc[1] += c[1]; b.tacc[2] += c[2];
}

/* ---------------- joint construction (ode/src/joints/*.cpp set-up functions) ---------------- */
inline void joint_set_fixed(Joint& j) { /* dJointSetFixed */
    dReal ofs[3] = {j.b0->pos[0] - j.b1->pos[0], j.b0->pos[1] - j.b1->pos[1], j.b0->pos[2] - j.b1->pos[2]};
    mul1_331(j.offset, j.b0->R, ofs);
    qmul1(j.qrel, j.b0->q, j.b1->q);
}
inline void joint_set_ball_anchor(Joint& j, const dReal* p) { /* dJointSetBallAnchor -> setAnchors */
    body_pos_rel_point(*j.b0, p, j.anchor1);
    body_pos_rel_point(*j.b1, p, j.anchor2);
}
inline void joint_set_slider_axis(Joint& j, const dReal* axis) { /* dJointSetSliderAxis */
    dReal a[3] = {axis[0], axis[1], axis[2]}; normalize3(a);
    mul1_331(j.axis1, j.b0->R, a);
    dReal c[3] = {j.b0->pos[0] - j.b1->pos[0], j.b0->pos[1] - j.b1->pos[1], j.b0->pos[2] - j.b1->pos[2]};
    mul1_331(j.offset, j.b1->R, c);
    qmul1(j.qrel, j.b0->q, j.b1->q);
}
inline dReal dball_current_distance(const Joint& j) {
    dReal g1[3], g2[3]; body_rel_point_pos(*j.b0, j.anchor1, g1); body_rel_point_pos(*j.b1, j.anchor2, g2);
    dReal d[3] = {g1[0] - g2[0], g1[1] - g2[1], g1[2] - g2[2]};
    return sqrtf(dot3(d, d));
}
inline void joint_set_dball_anchor1(Joint& j, const dReal* p) { body_pos_rel_point(*j.b0, p, j.anchor1); j.targetDistance = dball_current_distance(j); }
inline void joint_set_dball_anchor2(Joint& j, const dReal* p) { body_pos_rel_point(*j.b1, p, j.anchor2); j.targetDistance = dball_current_distance(j); }

/* ---------------- constraint rows ---------------- */
struct Row { dReal J1[6], J2[6], c, cfm; int b0, b1; dReal lo, hi; int findex; };   /* lo / hi / findex: contact rows only (bilateral rows: unbounded) */

/* contact.cpp getInfo2 for a contact whose second body is the static world: row 0 the normal (lo 0, soft ERP / CFM, bounce,
 * push-out capped by ContactMaxCorrectingVel), rows 1-2 the friction directions of dPlaneSpace(normal) with dContactApprox1
 * (bounds = +-mu x normal force, findex = row 0); returns 3 */
inline int contact_get_rows(const World& w, const ContactJoint& cj, dReal fps, Row* r, int firstRow) {
    const Body& b = *cj.b0;
    for (int i = 0; i < 3; ++i) { memset(&r[i], 0, sizeof(Row)); r[i].b0 = b.index; r[i].b1 = b.index; r[i].cfm = w.cfm; r[i].findex = -1; }
    dReal c1[3] = {cj.pos[0] - b.pos[0], cj.pos[1] - b.pos[1], cj.pos[2] - b.pos[2]};
    const dReal* n = cj.normal;
    for (int k = 0; k < 3; ++k) r[0].J1[k] = n[k];
    cross3(r[0].J1 + 3, c1, n);
    const dReal erp = cj.soft_erp >= 0 ? cj.soft_erp : w.erp;
    const dReal k = fps * erp;
    dReal depth = cj.depth - w.contactSurfaceLayer; if (depth < 0) depth = 0;
    r[0].cfm = cj.soft_cfm;
    dReal c = k * depth;
    if (c > w.contactMaxCorrectingVel) c = w.contactMaxCorrectingVel;
    {   /* bounce (bounce_vel = 0: the struct is zero-initialised by the reference) */
        const dReal outgoing = dot3(r[0].J1, b.lvel) + dot3(r[0].J1 + 3, b.avel);
        if (-outgoing > 0) { const dReal newc = -cj.bounce * outgoing; if (newc > c) c = newc; }
    }
    r[0].c = c; r[0].lo = 0; r[0].hi = 3.4e38f;
    dReal t1[3], t2[3]; plane_space(n, t1, t2);
    for (int k3 = 0; k3 < 3; ++k3) { r[1].J1[k3] = t1[k3]; r[2].J1[k3] = t2[k3]; }
    cross3(r[1].J1 + 3, c1, t1); cross3(r[2].J1 + 3, c1, t2);
    r[1].lo = -cj.mu; r[1].hi = cj.mu; r[1].findex = firstRow; r[2].lo = -cj.mu; r[2].hi = cj.mu; r[2].findex = firstRow;
    return 3;
}

inline void set_fixed_orientation(const Joint& j, dReal fps, dReal erp, Row* r) { /* joint.cpp setFixedOrientation */
    for (int i = 0; i < 3; ++i) { r[i].J1[3 + i] = 1; r[i].J2[3 + i] = -1; }
    dReal qq[4], qerr[4], e[3];
    qmul1(qq, j.b0->q, j.b1->q);
    qmul2(qerr, qq, j.qrel);
    if (qerr[0] < 0) { qerr[1] = -qerr[1]; qerr[2] = -qerr[2]; qerr[3] = -qerr[3]; }
    mul0_331(e, j.b0->R, qerr + 1);
    dReal k2 = fps * erp * 2.0f;
    for (int i = 0; i < 3; ++i) r[i].c = k2 * e[i];
}

inline int joint_get_rows(const World& w, const Joint& j, dReal fps, Row* r) {
    const int n = j.rows();
    for (int i = 0; i < n; ++i) {
        memset(&r[i], 0, sizeof(Row));
        r[i].cfm = w.cfm; r[i].b0 = j.b0->index; r[i].b1 = j.b1->index;
    }
    const Body& b0 = *j.b0; const Body& b1 = *j.b1;
    switch (j.type) {
    case J_BALL: { /* ball.cpp getInfo2 -> joint.cpp setBall */
        dReal a1[3], a2[3];
        mul0_331(a1, b0.R, j.anchor1); mul0_331(a2, b1.R, j.anchor2);
        for (int i = 0; i < 3; ++i) { r[i].J1[i] = 1; r[i].J2[i] = -1; r[i].cfm = j.cfm; }
        /* J1a = -[a1]x , J2a = +[a2]x */
        r[0].J1[4] = a1[2];  r[0].J1[5] = -a1[1]; r[1].J1[3] = -a1[2]; r[1].J1[5] = a1[0]; r[2].J1[3] = a1[1];  r[2].J1[4] = -a1[0];
        r[0].J2[4] = -a2[2]; r[0].J2[5] = a2[1];  r[1].J2[3] = a2[2];  r[1].J2[5] = -a2[0]; r[2].J2[3] = -a2[1]; r[2].J2[4] = a2[0];
        dReal k = fps * j.erp;
        for (int i = 0; i < 3; ++i) r[i].c = k * (a2[i] + b1.pos[i] - a1[i] - b0.pos[i]);
        break; }
    case J_FIXED: { /* fixed.cpp getInfo2: rows 0-2 linear, 3-5 angular */
        set_fixed_orientation(j, fps, j.erp, r + 3);
        dReal ofs[3]; mul0_331(ofs, b0.R, j.offset);
        for (int i = 0; i < 3; ++i) { r[i].J1[i] = 1; r[i].J2[i] = -1; }
        /* J1a = +[ofs]x */
        r[0].J1[4] = -ofs[2]; r[0].J1[5] = ofs[1]; r[1].J1[3] = ofs[2]; r[1].J1[5] = -ofs[0]; r[2].J1[3] = -ofs[1]; r[2].J1[4] = ofs[0];
        dReal k = fps * j.erp;
        for (int i = 0; i < 3; ++i) r[i].c = k * (b1.pos[i] - b0.pos[i] + ofs[i]);
        for (int i = 0; i < 6; ++i) r[i].cfm = j.cfm;
        break; }
    case J_SLIDER: { /* slider.cpp getInfo2 (no limits, no motor -> 5 rows) */
        set_fixed_orientation(j, fps, w.erp, r);
        dReal c[3] = {b1.pos[0] - b0.pos[0], b1.pos[1] - b0.pos[1], b1.pos[2] - b0.pos[2]};
        dReal ax1[3], p[3], q[3];
        mul0_331(ax1, b0.R, j.axis1);
        plane_space(ax1, p, q);
        dReal t[3];
        for (int i = 0; i < 3; ++i) { r[3].J1[i] = p[i]; r[4].J1[i] = q[i]; r[3].J2[i] = -p[i]; r[4].J2[i] = -q[i]; }
        cross3(t, c, p); for (int i = 0; i < 3; ++i) { r[3].J1[3 + i] = 0.5f * t[i]; r[3].J2[3 + i] = 0.5f * t[i]; }
        cross3(t, c, q); for (int i = 0; i < 3; ++i) { r[4].J1[3 + i] = 0.5f * t[i]; r[4].J2[3 + i] = 0.5f * t[i]; }
        dReal ofs[3]; mul0_331(ofs, b1.R, j.offset);
        for (int i = 0; i < 3; ++i) c[i] += ofs[i];
        dReal k = fps * w.erp;
        r[3].c = k * dot3(p, c); r[4].c = k * dot3(q, c);
        break; }
    case J_DBALL: { /* dball.cpp getInfo2 */
        dReal g1[3], g2[3], q[3];
        body_rel_point_pos(b0, j.anchor1, g1); body_rel_point_pos(b1, j.anchor2, g2);
        for (int i = 0; i < 3; ++i) q[i] = g1[i] - g2[i];
        dReal dist = sqrtf(dot3(q, q));
        if (dist < 1e-7f) {
            dReal v1[3], v2[3]; body_point_vel(b0, g1, v1); body_point_vel(b1, g2, v2);
            for (int i = 0; i < 3; ++i) q[i] = v1[i] - v2[i];
            if (sqrtf(dot3(q, q)) < 1e-7f) { q[0] = 1; q[1] = 0; q[2] = 0; }
        }
        normalize3(q);
        dReal ra1[3], ra2[3], t[3];
        mul0_331(ra1, b0.R, j.anchor1); mul0_331(ra2, b1.R, j.anchor2);
        for (int i = 0; i < 3; ++i) { r[0].J1[i] = q[i]; r[0].J2[i] = -q[i]; }
        cross3(t, ra1, q); for (int i = 0; i < 3; ++i) r[0].J1[3 + i] = t[i];
        cross3(t, q, ra2); for (int i = 0; i < 3; ++i) r[0].J2[3 + i] = t[i];
        r[0].c = fps * j.erp * (j.targetDistance - dist);
        r[0].cfm = j.cfm;
        break; }
    }
    return n;
}

/* 3x3 helpers for world inertia */
inline void world_inertia(const Body& b, dReal* Iw, dReal* invIw) {
    /* I_w = R diag(I) R^T */
    for (int i = 0; i < 3; ++i)
        for (int k = 0; k < 3; ++k) {
            dReal s = 0, si = 0;
            for (int a = 0; a < 3; ++a) { s += b.R[i * 3 + a] * b.I[a] * b.R[k * 3 + a]; si += b.R[i * 3 + a] * (1.0f / b.I[a]) * b.R[k * 3 + a]; }
            Iw[i * 3 + k] = s; invIw[i * 3 + k] = si;
        }
}
inline void mat3_mul_vec(dReal* r, const dReal* M, const dReal* v) { r[0] = dot3(M, v); r[1] = dot3(M + 3, v); r[2] = dot3(M + 6, v); }
inline bool mat3_invert(dReal* o, const dReal* m) {
    dReal c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    dReal det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    if (det == 0) return false;
    dReal id = 1.0f / det;
    o[0] = c00 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    o[3] = c01 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    o[6] = c02 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
    return true;
}
inline dReal sinc_ode(dReal x) { return fabsf(x) < 1.0e-4f ? 1.0f - x * x * 0.166666666666666666667f : sinf(x) / x; }

/* util.cpp dxStepBody with dxBodyFlagFiniteRotation set and no finite-rotation axis
 * (RigidBodyODE.cpp:15-16: mode 1, axis (0,0,0)); no damping, no max angular speed. */
inline void step_body(Body& b, dReal h) {
    for (int j = 0; j < 3; ++j) b.pos[j] += h * b.lvel[j];
    dReal wlen = sqrtf(b.avel[0] * b.avel[0] + b.avel[1] * b.avel[1] + b.avel[2] * b.avel[2]);
    h *= 0.5f;
    dReal theta = wlen * h;
    dReal q[4]; q[0] = cosf(theta);
    dReal s = sinc_ode(theta) * h;
    q[1] = b.avel[0] * s; q[2] = b.avel[1] * s; q[3] = b.avel[2] * s;
    dReal q2[4]; qmul0(q2, q, b.q);
    for (int j = 0; j < 4; ++j) b.q[j] = q2[j];
    normalize4(b.q);
    q_to_R(b.q, b.R);
}

/* step.cpp dxStepIsland for one island holding every body/joint of the world (one car per world:
 * each car is its own island).  All rows unbounded -> dSolveLCP degenerates to one
 * LDL^T factor + solve. */
inline void world_step(World& w, dReal h) {
    const int nb = (int)w.bodies.size();
    const dReal hinv = 1.0f / h;
    std::vector<dReal> Iw(nb * 9), invIw(nb * 9);
    for (int i = 0; i < nb; ++i) {
        Body& b = *w.bodies[i]; b.index = i;
        world_inertia(b, &Iw[i * 9], &invIw[i * 9]);
        /* gyroscopic term, implicit form (step.cpp, dxBodyGyroscopic is set by dBodyCreate) */
        {
            const dReal* I = &Iw[i * 9];
            dReal L[3]; mat3_mul_vec(L, I, b.avel);
            dReal It[9] = {0, L[2], -L[1], -L[2], 0, L[0], L[1], -L[0], 0}; /* dSetCrossMatrixMinus */
            for (int k = 0; k < 9; ++k) It[k] = It[k] * h + I[k];
            dReal Ls[3] = {L[0] * hinv, L[1] * hinv, L[2] * hinv};
            dReal inv[9];
            if (mat3_invert(inv, It)) {
                dReal M[9];
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r * 3 + c] = I[r * 3 + 0] * inv[0 * 3 + c] + I[r * 3 + 1] * inv[1 * 3 + c] + I[r * 3 + 2] * inv[2 * 3 + c];
                M[0] -= 1; M[4] -= 1; M[8] -= 1;
                dReal tau[3]; mat3_mul_vec(tau, M, Ls);
                b.tacc[0] += tau[0]; b.tacc[1] += tau[1]; b.tacc[2] += tau[2];
            }
        }
        for (int k = 0; k < 3; ++k) b.facc[k] += b.mass * w.gravity[k];
    }
    /* rows */
    int mb = 0; for (Joint* j : w.joints) mb += j->rows();           /* bilateral (unbounded) rows first, as dSolveLCP orders them */
    const int mc = 3 * (int)w.contacts.size();
    const int m = mb + mc;
    std::vector<Row> rows(m);
    { int o = 0; for (Joint* j : w.joints) { const int n = joint_get_rows(w, *j, hinv, &rows[o]); for (int i = 0; i < n; ++i) { rows[o + i].lo = -3.4e38f; rows[o + i].hi = 3.4e38f; rows[o + i].findex = -1; } o += n; }
      for (const ContactJoint& cj : w.contacts) o += contact_get_rows(w, cj, hinv, &rows[o], o); }
    w.last_m = m;
    std::vector<dReal> lambda(m, 0.0f);
    if (m > 0) {
        /* JinvM */
        std::vector<dReal> JiM(m * 12);
        for (int i = 0; i < m; ++i) {
            const Row& r = rows[i];
            const Body& b0 = *w.bodies[r.b0]; const Body& b1 = *w.bodies[r.b1];
            dReal im0 = 1.0f / b0.mass, im1 = 1.0f / b1.mass;
            for (int k = 0; k < 3; ++k) { JiM[i * 12 + k] = r.J1[k] * im0; JiM[i * 12 + 6 + k] = r.J2[k] * im1; }
            mat3_mul_vec(&JiM[i * 12 + 3], &invIw[r.b0 * 9], r.J1 + 3);
            mat3_mul_vec(&JiM[i * 12 + 9], &invIw[r.b1 * 9], r.J2 + 3);
        }
        /* A = JinvM * J^T (lower triangle), + cfm/h on the diagonal */
        std::vector<dReal> A((size_t)m * m, 0.0f);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j <= i; ++j) {
                const Row& ri = rows[i]; const Row& rj = rows[j];
                dReal s = 0;
                const dReal* a0 = &JiM[i * 12]; const dReal* a1 = &JiM[i * 12 + 6];
                if (ri.b0 == rj.b0) for (int k = 0; k < 6; ++k) s += a0[k] * rj.J1[k];
                if (ri.b0 == rj.b1 && rj.b1 != rj.b0) for (int k = 0; k < 6; ++k) s += a0[k] * rj.J2[k];
                if (ri.b1 == rj.b0 && ri.b1 != ri.b0) for (int k = 0; k < 6; ++k) s += a1[k] * rj.J1[k];
                if (ri.b1 == rj.b1 && ri.b1 != ri.b0 && rj.b1 != rj.b0) for (int k = 0; k < 6; ++k) s += a1[k] * rj.J2[k];
                A[(size_t)i * m + j] = s;
            }
        for (int i = 0; i < m; ++i) A[(size_t)i * m + i] += rows[i].cfm * hinv;
        /* rhs = c/h - J*(v/h + invM*fe) */
        std::vector<dReal> tmp1(nb * 6);
        for (int i = 0; i < nb; ++i) {
            const Body& b = *w.bodies[i]; dReal im = 1.0f / b.mass;
            for (int k = 0; k < 3; ++k) tmp1[i * 6 + k] = b.facc[k] * im + b.lvel[k] * hinv;
            dReal t[3]; mat3_mul_vec(t, &invIw[i * 9], b.tacc);
            for (int k = 0; k < 3; ++k) tmp1[i * 6 + 3 + k] = t[k] + b.avel[k] * hinv;
        }
        std::vector<dReal> rhs(m);
        for (int i = 0; i < m; ++i) {
            const Row& r = rows[i]; dReal s = 0;
            for (int k = 0; k < 6; ++k) s += r.J1[k] * tmp1[r.b0 * 6 + k];
            if (r.b1 != r.b0) for (int k = 0; k < 6; ++k) s += r.J2[k] * tmp1[r.b1 * 6 + k];
            rhs[i] = r.c * hinv - s;
        }
        /* LDL^T factor + solve (fastldlt.c semantics: unit-lower L and diagonal D, row by row:
           solve L(0:i,0:i) u = A(i,0:i), then L_ij = u_j / D_j and D_i = A_ii - sum_j u_j L_ij) */
        /* with contact rows: dSolveLCP factors the unbounded block first; the bounded rows then see the Schur complement
           A_cc - A_cb A_bb^-1 A_bc.  ODE continues with Dantzig pivoting; here the small bounded system is solved by projected
           Gauss-Seidel (a fixed number of sweeps in row order, friction bounds +-mu x the current normal force), see the header
           of ode_collide.h: the contact path is this repository's own definition. */
        if (w.keepSystem) { w.lastA.assign((size_t)mb * mb, 0.0f); for (int i = 0; i < mb; ++i) for (int j = 0; j <= i; ++j) { w.lastA[(size_t)i * mb + j] = A[(size_t)i * m + j]; w.lastA[(size_t)j * mb + i] = A[(size_t)i * m + j]; } w.lastRhs.assign(rhs.begin(), rhs.begin() + mb); }
        std::vector<dReal> Acc, rc;
        if (mc > 0) { Acc.assign((size_t)mc * mc, 0.0f); rc.assign(mc, 0.0f);
            for (int i = 0; i < mc; ++i) { for (int j = 0; j <= i; ++j) { Acc[(size_t)i * mc + j] = A[(size_t)(mb + i) * m + mb + j]; Acc[(size_t)j * mc + i] = Acc[(size_t)i * mc + j]; } rc[i] = rhs[mb + i]; } }
        std::vector<dReal> Acb;
        if (mc > 0) { Acb.assign((size_t)mc * mb, 0.0f); for (int i = 0; i < mc; ++i) for (int j = 0; j < mb; ++j) Acb[(size_t)i * mb + j] = A[(size_t)(mb + i) * m + j]; }
        const int mfull = m; (void)mfull;
        std::vector<dReal> d(mb);
        for (int i = 0; i < mb; ++i) {
            dReal* Ai = &A[(size_t)i * m];
            for (int j = 0; j < i; ++j) {
                const dReal* Aj = &A[(size_t)j * m];
                dReal s = Ai[j];
                for (int k = 0; k < j; ++k) s -= Ai[k] * Aj[k];   /* Ai[k] = u_k (unscaled), Aj[k] = L_jk */
                Ai[j] = s;
            }
            dReal dii = Ai[i];
            for (int j = 0; j < i; ++j) {
                dReal lij = Ai[j] / d[j];
                dii -= Ai[j] * lij;
                Ai[j] = lij;
            }
            d[i] = dii;
        }
        auto solve_b = [&](std::vector<dReal>& x) {          /* x <- A_bb^-1 x with the factor above */
            for (int i = 0; i < mb; ++i) { dReal s = x[i]; for (int k = 0; k < i; ++k) s -= A[(size_t)i * m + k] * x[k]; x[i] = s; }
            for (int i = 0; i < mb; ++i) x[i] /= d[i];
            for (int i = mb - 1; i >= 0; --i) { dReal s = x[i]; for (int k = i + 1; k < mb; ++k) s -= A[(size_t)k * m + i] * x[k]; x[i] = s; }
        };
        std::vector<dReal> xb(rhs.begin(), rhs.begin() + mb);
        solve_b(xb);                                         /* A_bb^-1 rhs_b */
        if (mc > 0) {
            /* Schur complement onto the contact rows */
            std::vector<dReal> col(mb);
            for (int j = 0; j < mc; ++j) {
                for (int k = 0; k < mb; ++k) col[k] = Acb[(size_t)j * mb + k];     /* A_bc(:, j) = A_cb(j, :)^T */
                solve_b(col);
                for (int i = 0; i < mc; ++i) { dReal s = 0; for (int k = 0; k < mb; ++k) s += Acb[(size_t)i * mb + k] * col[k]; Acc[(size_t)i * mc + j] -= s; }
            }
            for (int i = 0; i < mc; ++i) { dReal s = 0; for (int k = 0; k < mb; ++k) s += Acb[(size_t)i * mb + k] * xb[k]; rc[i] -= s; }
            std::vector<dReal> lc(mc, 0.0f);
            for (int it = 0; it < w.pgsIterations; ++it)
                for (int i = 0; i < mc; ++i) {
                    const Row& r = rows[mb + i];
                    dReal s = rc[i]; for (int k = 0; k < mc; ++k) s -= Acc[(size_t)i * mc + k] * lc[k];
                    dReal v = lc[i] + s / Acc[(size_t)i * mc + i];
                    dReal lo = r.lo, hi = r.hi;
                    if (r.findex >= 0) { const dReal fn = lc[r.findex - mb]; hi = r.hi * fn; lo = -hi; }
                    if (v < lo) v = lo; if (v > hi) v = hi;
                    lc[i] = v;
                }
            /* lambda_b = A_bb^-1 (rhs_b - A_bc lambda_c) */
            for (int k = 0; k < mb; ++k) { dReal s = rhs[k]; for (int i = 0; i < mc; ++i) s -= Acb[(size_t)i * mb + k] * lc[i]; xb[k] = s; }
            solve_b(xb);
            for (int i = 0; i < mc; ++i) lambda[mb + i] = lc[i];
        }
        for (int i = 0; i < mb; ++i) lambda[i] = xb[i];
    }
    w.last_lambda = lambda;
    /* cforce = J^T lambda ; velocity update ; position update */
    std::vector<dReal> cf(nb * 6, 0.0f);
    for (int i = 0; i < m; ++i) {
        const Row& r = rows[i];
        for (int k = 0; k < 6; ++k) { cf[r.b0 * 6 + k] += r.J1[k] * lambda[i]; if (r.b1 != r.b0) cf[r.b1 * 6 + k] += r.J2[k] * lambda[i]; }
    }
    for (int i = 0; i < nb; ++i) {
        Body& b = *w.bodies[i];
        dReal imh = h / b.mass;
        for (int k = 0; k < 3; ++k) b.lvel[k] += (cf[i * 6 + k] + b.facc[k]) * imh;
        dReal t[3] = {(cf[i * 6 + 3] + b.tacc[0]) * h, (cf[i * 6 + 4] + b.tacc[1]) * h, (cf[i * 6 + 5] + b.tacc[2]) * h};
        dReal dw[3]; mat3_mul_vec(dw, &invIw[i * 9], t);
        for (int k = 0; k < 3; ++k) b.avel[k] += dw[k];
        step_body(b, h);
        for (int k = 0; k < 3; ++k) { b.facc[k] = 0; b.tacc[k] = 0; }
    }
}

} // namespace oder
