/*
 * pd_contacts.h -- collision RESPONSE (second half): contact generation for the chassis colliders against the
 * static track meshes, and the contact joints' rows inside the constraint solve.
 *
 * Reference: PhysicsEngineODE::onCollision (Physics/ODE/PhysicsEngineODE.cpp:284-341) creates one dxJointContact per contact
 * ODE's colliders return (box <-> trimesh: mode 28700 = Bounce | SoftERP | SoftCFM | Approx1, mu 0.1, soft_cfm 9.5238e-4, soft_erp
 * 0.714286; anything else: mode 28692 = Bounce | SoftCFM | Approx1, mu 0.25, bounce 0.01, soft_cfm 1e-4) in the contact group of the
 * frame's parity, which is emptied only on its own parity (:228-244): a contact made on an odd frame acts in that frame's
 * dWorldStep and in the next one.  ODE's contact GENERATORS are absent from the reference tree and unpinned by it; the contact
 * set is therefore the one DEFINED in oracle/ode_restate/ode_collide.h (header), followed here to the letter:
 *   floor box vs accepted TRACK triangle -> box corners below the triangle's plane and inside its prism (per corner the deepest);
 *     none at all -> one contact at the support corner of the deepest accepted triangle's SAT normal;
 *   hull vs WALL triangle crossed by a hull triangle -> hull vertices less than 0.5 m behind its plane, inside its prism (per
 *     vertex the deepest); at most 4 + 4 contacts per car, deepest first.
 * Rows (ode/src/joints/contact.cpp getInfo2, second body = the static world): normal row (lo 0; c = min(erp fps depth,
 * ContactMaxCorrectingVel 3) or the bounce velocity if larger; cfm = soft_cfm), two friction rows along dPlaneSpace(normal) with
 * bounds +-mu x the normal force (dContactApprox1).
 *
 * Every contact row touches the chassis only, so with S z = b the chassis system of the block elimination (pd_solver.h) the rows
 * reduce to a 3k x 3k bounded system in the contact forces, (J S^-1 J^T + cfm / h) lambda = rhs - J S^-1 b, solved by projected
 * Gauss-Seidel (50 sweeps in row order, as the oracle does on the Schur complement of its dense system), then z = S^-1 (b + J^T lambda)
 * and the joint groups are back-substituted as always.  Only envs that HAVE live contacts take this path.
 */
#pragma once
#include "pd_collide.h"
#include "pd_solver2.h"

namespace pd {

#define PD_MAX_CONTACTS 8
#define PD_CONTACT_WORDS 68          /* per env: count | 8 x {pos3, normal3, depth, kind} | pad */
#define PD_CONTACT_HULL_SLOTS PD_MAX_COLLIDER_VERTS
#define PD_CONTACT_SLOTS (8 + PD_CONTACT_HULL_SLOTS + 1)          /* 8 box corners | one per hull vertex | the floor fallback */
#define PD_CONTACT_FALLBACK (PD_CONTACT_SLOTS - 1)

struct ContactSlot { float depth, nx, ny, nz, px, py, pz; };
/* total order of ode_collide.h contact_offer: deeper wins, ties by larger normal.y, then .x, then .z */
PD_HD bool contact_better(float depth, V3 n, const ContactSlot& c) {
    if (depth > c.depth) return true;
    if (depth < c.depth) return false;
    if (n.y > c.ny) return true;
    if (n.y < c.ny) return false;
    if (n.x > c.nx) return true;
    if (n.x < c.nx) return false;
    return n.z > c.nz;
}
PD_HD void contact_offer(ContactSlot& c, V3 pos, V3 n, float depth) {
    if (c.depth < 0.0f || contact_better(depth, n, c)) { c.depth = depth; c.nx = n.x; c.ny = n.y; c.nz = n.z; c.px = pos.x; c.py = pos.y; c.pz = pos.z; }
}
PD_HD bool projects_inside(V3 p, V3 v0, V3 v1, V3 v2, V3 N) {
    const float d0 = dot(cross(v1 - v0, p - v0), N);
    const float d1 = dot(cross(v2 - v1, p - v1), N);
    const float d2 = dot(cross(v0 - v2, p - v2), N);
    return d0 >= 0.0f && d1 >= 0.0f && d2 >= 0.0f;
}

#if defined(__CUDACC__)
/* Contact candidates of one car, one warp: the lanes share the cells / list entries of the footprint exactly as k_collide's
 * detection does (lane = cellPart * 8 + part), every lane keeps its own best candidate per slot in local memory, the slots are
 * then merged over the warp with the total order above (so the result does not depend on who saw which triangle) and lane 0
 * writes the selection.  Runs only for cars whose detection answered "contact". */
__device__ __noinline__ void car_contacts_warp(const PdCarParams& P, const TrackDev& T, const Body& C, int lane, float* out /* PD_CONTACT_WORDS */) {
    const unsigned FULL = 0xffffffffu;
    const PdBoundGrid& G = T.collGrid;
    ContactSlot slot[PD_CONTACT_SLOTS];
    for (int k = 0; k < PD_CONTACT_SLOTS; ++k) { slot[k].depth = -1.0f; slot[k].nx = slot[k].ny = slot[k].nz = 0.0f; slot[k].px = slot[k].py = slot[k].pz = 0.0f; }
    const Frame& f = C.fr;
    const bool hasBox = P.hasBoxCollider != 0, hasHull = P.nColliderTris > 0;
    const V3 bc = to_world(f, v3(P.boxCentre[0], P.boxCentre[1], P.boxCentre[2]));
    const V3 bh = v3(P.boxSize[0] * 0.5f, P.boxSize[1] * 0.5f, P.boxSize[2] * 0.5f);
    V3 blo, bhi; obb_bounds(f, bc, bh, blo, bhi);
    const V3 hcl = v3((P.colliderMin[0] + P.colliderMax[0]) * 0.5f, (P.colliderMin[1] + P.colliderMax[1]) * 0.5f, (P.colliderMin[2] + P.colliderMax[2]) * 0.5f);
    const V3 hh = v3((P.colliderMax[0] - P.colliderMin[0]) * 0.5f + 1e-3f, (P.colliderMax[1] - P.colliderMin[1]) * 0.5f + 1e-3f, (P.colliderMax[2] - P.colliderMin[2]) * 0.5f + 1e-3f);
    const V3 hc = to_world(f, hcl);
    V3 hlo, hhi; obb_bounds(f, hc, hh, hlo, hhi);
    float x0 = hasBox ? blo.x : hlo.x, x1 = hasBox ? bhi.x : hhi.x, z0 = hasBox ? blo.z : hlo.z, z1 = hasBox ? bhi.z : hhi.z;
    if (hasHull) { x0 = tminf(x0, hlo.x); x1 = tmaxf(x1, hhi.x); z0 = tminf(z0, hlo.z); z1 = tmaxf(z1, hhi.z); }
    int ix0 = (int)floorf((x0 - G.ox) * G.invCell), ix1 = (int)floorf((x1 - G.ox) * G.invCell);
    int iz0 = (int)floorf((z0 - G.oz) * G.invCell), iz1 = (int)floorf((z1 - G.oz) * G.invCell);
    if (ix0 < 0) ix0 = 0; if (iz0 < 0) iz0 = 0; if (ix1 >= G.nx) ix1 = G.nx - 1; if (iz1 >= G.nz) iz1 = G.nz - 1;
    if (ix1 - ix0 > 8) ix1 = ix0 + 8;
    if (iz1 - iz0 > 8) iz1 = iz0 + 8;
    const V3 cmin = v3(P.colliderMin[0] - 1e-4f, P.colliderMin[1] - 1e-4f, P.colliderMin[2] - 1e-4f), cmax = v3(P.colliderMax[0] + 1e-4f, P.colliderMax[1] + 1e-4f, P.colliderMax[2] + 1e-4f);
    const int wx = ix1 - ix0 + 1, ncell = (G.nx > 0 && G.nz > 0) ? wx * (iz1 - iz0 + 1) : 0;
    const int cellPart = lane >> 3, part = lane & 7;
    V3 corner[8];
    for (int k = 0; k < 8; ++k)
        corner[k] = to_world(f, v3(P.boxCentre[0] + ((k & 1) ? bh.x : -bh.x), P.boxCentre[1] + ((k & 2) ? bh.y : -bh.y), P.boxCentre[2] + ((k & 4) ? bh.z : -bh.z)));
    for (int ci = cellPart; ci < ncell; ci += 4) {
        const int c = (iz0 + ci / wx) * G.nx + (ix0 + ci % wx);
        const float4 yr = __ldg(reinterpret_cast<const float4*>(T.collCell) + 2 * (size_t)c), kk = __ldg(reinterpret_cast<const float4*>(T.collCell) + 2 * (size_t)c + 1);
        const float tY0 = yr.x, tY1 = yr.y, wY0 = yr.z, wY1 = yr.w;
        const int kT = __float_as_int(kk.x), kW = __float_as_int(kk.y), kE = __float_as_int(kk.z);
        if (hasBox && !(tY0 > bhi.y || tY1 < blo.y)) {
            for (int k = kT + part; k < kW; k += 8) {
                float mn[3], mx[3]; int tri; load_coll_rec(T.collRec, k, mn, mx, tri);
                if (mx[1] < blo.y) break;
                if (mn[1] > bhi.y || mn[0] > bhi.x || mx[0] < blo.x || mn[2] > bhi.z || mx[2] < blo.z) continue;
                const float* q = T.triRaw + (size_t)tri * 9;
                const V3 t0 = v3(q[0], q[1], q[2]), t1 = v3(q[3], q[4], q[5]), t2 = v3(q[6], q[7], q[8]);
                V3 n; float satDepth = 0.0f;
                if (!box_tri_contact(bc, f.ax, f.ay, f.az, bh, t0, t1, t2, n, &satDepth)) continue;
                if (dot(f.ay, n) < 0.9f) continue;
                contact_offer(slot[PD_CONTACT_FALLBACK], v3(0, 0, 0), n, satDepth);                      /* fallback: the deepest accepted triangle */
                const V3 N = cross(t1 - t0, t2 - t0);
                const float len = sqrtf(dot(N, N));
                if (!(len > 1e-12f)) continue;
                const float inv = 1.0f / len;
                V3 nt = v3(N.x * inv, N.y * inv, N.z * inv);
                if (dot(nt, n) < 0.0f) nt = v3(-nt.x, -nt.y, -nt.z);
                for (int cz = 0; cz < 8; ++cz) {
                    const float sdist = dot(corner[cz] - t0, nt);
                    if (!(sdist < 0.0f)) continue;
                    if (!projects_inside(corner[cz], t0, t1, t2, N)) continue;
                    contact_offer(slot[cz], corner[cz], nt, -sdist);
                }
            }
        }
        if (hasHull && !(wY0 > hhi.y || wY1 < hlo.y)) {
            for (int k = kW + part; k < kE; k += 8) {
                float mn[3], mx[3]; int tri; load_coll_rec(T.collRec, k, mn, mx, tri);
                if (mx[1] < hlo.y) break;
                if (mn[1] > hhi.y || mn[0] > hhi.x || mx[0] < hlo.x || mn[2] > hhi.z || mx[2] < hlo.z) continue;
                const float* q = T.triRaw + (size_t)tri * 9;
                const V3 b0 = to_local(f, v3(q[0], q[1], q[2])), b1 = to_local(f, v3(q[3], q[4], q[5])), b2 = to_local(f, v3(q[6], q[7], q[8]));
                const V3 lo = v3(fminf(b0.x, fminf(b1.x, b2.x)), fminf(b0.y, fminf(b1.y, b2.y)), fminf(b0.z, fminf(b1.z, b2.z)));
                const V3 hi = v3(fmaxf(b0.x, fmaxf(b1.x, b2.x)), fmaxf(b0.y, fmaxf(b1.y, b2.y)), fmaxf(b0.z, fmaxf(b1.z, b2.z)));
                if (lo.x > cmax.x || hi.x < cmin.x || lo.y > cmax.y || hi.y < cmin.y || lo.z > cmax.z || hi.z < cmin.z) continue;
                bool triHit = false;
                for (int j = 0; j < P.nColliderTris && !triHit; ++j) {
                    const float* tb = P.colliderTriBounds[j];
                    if (lo.x > tb[3] || hi.x < tb[0] || lo.y > tb[4] || hi.y < tb[1] || lo.z > tb[5] || hi.z < tb[2]) continue;
                    const float* p0 = P.colliderVerts[P.colliderTris[j][0]]; const float* p1 = P.colliderVerts[P.colliderTris[j][1]]; const float* p2 = P.colliderVerts[P.colliderTris[j][2]];
                    if (tri_tri(v3(p0[0], p0[1], p0[2]), v3(p1[0], p1[1], p1[2]), v3(p2[0], p2[1], p2[2]), b0, b1, b2)) triHit = true;
                }
                if (!triHit) continue;
                const V3 N = cross(b1 - b0, b2 - b0);
                const float len = sqrtf(dot(N, N));
                if (!(len > 1e-12f)) continue;
                const float inv = 1.0f / len;
                V3 nl = v3(N.x * inv, N.y * inv, N.z * inv);
                if (dot(v3(0, 0, 0) - b0, nl) < 0.0f) nl = v3(-nl.x, -nl.y, -nl.z);
                const V3 nw = rot(f, nl);
                for (int j = 0; j < P.nColliderVerts && j < PD_CONTACT_HULL_SLOTS; ++j) {
                    const V3 hv = v3(P.colliderVerts[j][0], P.colliderVerts[j][1], P.colliderVerts[j][2]);
                    const float sdist = dot(hv - b0, nl);
                    if (!(sdist < 0.0f) || !(sdist > -0.5f)) continue;
                    if (!projects_inside(hv, b0, b1, b2, N)) continue;
                    contact_offer(slot[8 + j], to_world(f, hv), nw, -sdist);
                }
            }
        }
    }
    /* merge over the warp: butterfly with the total order (identical candidates from several lanes are the same value) */
    for (int k = 0; k < PD_CONTACT_SLOTS; ++k) {
        ContactSlot s = slot[k];
        for (int off = 16; off > 0; off >>= 1) {
            ContactSlot o;
            o.depth = __shfl_xor_sync(FULL, s.depth, off); o.nx = __shfl_xor_sync(FULL, s.nx, off); o.ny = __shfl_xor_sync(FULL, s.ny, off); o.nz = __shfl_xor_sync(FULL, s.nz, off);
            o.px = __shfl_xor_sync(FULL, s.px, off); o.py = __shfl_xor_sync(FULL, s.py, off); o.pz = __shfl_xor_sync(FULL, s.pz, off);
            if (o.depth >= 0.0f && (s.depth < 0.0f || contact_better(o.depth, v3(o.nx, o.ny, o.nz), s))) s = o;
        }
        slot[k] = s;
    }
    if (lane != 0) return;
    int n = 0;
    {   /* floor: the 4 deepest corners, ties by lower corner index; none -> the fallback contact */
        bool taken[8] = {false, false, false, false, false, false, false, false};
        int nf = 0;
        for (int r = 0; r < 4; ++r) {
            int best = -1;
            for (int k = 0; k < 8; ++k) if (slot[k].depth >= 0.0f && !taken[k] && (best < 0 || slot[k].depth > slot[best].depth)) best = k;
            if (best < 0) break;
            taken[best] = true; ++nf;
            float* o = out + 1 + n * 8; o[0] = slot[best].px; o[1] = slot[best].py; o[2] = slot[best].pz; o[3] = slot[best].nx; o[4] = slot[best].ny; o[5] = slot[best].nz; o[6] = slot[best].depth; o[7] = 0.0f; ++n;
        }
        if (nf == 0 && slot[PD_CONTACT_FALLBACK].depth >= 0.0f) {
            const V3 fbN = v3(slot[PD_CONTACT_FALLBACK].nx, slot[PD_CONTACT_FALLBACK].ny, slot[PD_CONTACT_FALLBACK].nz);
            int bestK = 0; float bestS = 3.4e38f;
            for (int k = 0; k < 8; ++k) { const float sdot = dot(corner[k], fbN); if (sdot < bestS) { bestS = sdot; bestK = k; } }
            float* o = out + 1 + n * 8; o[0] = corner[bestK].x; o[1] = corner[bestK].y; o[2] = corner[bestK].z; o[3] = fbN.x; o[4] = fbN.y; o[5] = fbN.z; o[6] = slot[PD_CONTACT_FALLBACK].depth; o[7] = 0.0f; ++n;
        }
    }
    {   /* walls: the 4 deepest hull vertices, ties by lower vertex index */
        bool taken[PD_CONTACT_HULL_SLOTS];
        for (int k = 0; k < PD_CONTACT_HULL_SLOTS; ++k) taken[k] = false;
        for (int r = 0; r < 4; ++r) {
            int best = -1;
            for (int k = 0; k < PD_CONTACT_HULL_SLOTS; ++k) if (slot[8 + k].depth >= 0.0f && !taken[k] && (best < 0 || slot[8 + k].depth > slot[8 + best].depth)) best = k;
            if (best < 0) break;
            taken[best] = true;
            const ContactSlot& s = slot[8 + best];
            float* o = out + 1 + n * 8; o[0] = s.px; o[1] = s.py; o[2] = s.pz; o[3] = s.nx; o[4] = s.ny; o[5] = s.nz; o[6] = s.depth; o[7] = 1.0f; ++n;
        }
    }
    reinterpret_cast<int*>(out)[0] = n;
}
#endif

/* The contact rows on top of the chassis system S z = b (S packed lower 6x6 with the chassis mass matrix already added):
 * returns z including the contact forces.  Also applies the side effect of Car::onCollisionCallback that lives in the state
 * record -- the engine blows up (lifeLeft = -100) when a wall is met at more than 150 km/h x mechanicalDamageRate, the damage zone
 * of the impact point takes the impact speed (Car.cpp:975-999, Engine.cpp:406-409) -- on the frame that created the contacts. */
PD_HDN void contacts_solve(const PdCarParams& P, const float* __restrict__ cont, const Body& C, const BodyDyn& dC, const float* __restrict__ S21, const float* __restrict__ b6, float h, bool fresh, float& lifeLeft, float* __restrict__ dmg, float* __restrict__ z) {
    const float hinv = 1.0f / h;
    int n = reinterpret_cast<const int*>(cont)[0];
    if (n > PD_MAX_CONTACTS) n = PD_MAX_CONTACTS;
    const int m = 3 * n;
    float J[3 * PD_MAX_CONTACTS][6], W[3 * PD_MAX_CONTACTS][6], rc[3 * PD_MAX_CONTACTS], cfm[3 * PD_MAX_CONTACTS], mu[PD_MAX_CONTACTS], lam[3 * PD_MAX_CONTACTS];
    float A[3 * PD_MAX_CONTACTS][3 * PD_MAX_CONTACTS];
    float z0[6];
    solve6(S21, b6, z0);
    for (int i = 0; i < n; ++i) {
        const float* c = cont + 1 + i * 8;
        const V3 pos = v3(c[0], c[1], c[2]), nrm = v3(c[3], c[4], c[5]);
        const float depthIn = c[6]; const bool wall = c[7] != 0.0f;
        const V3 r = pos - C.fr.p;
        const float erp = wall ? P.worldERP : 0.714285731f;
        float depth = depthIn; if (depth < 0.0f) depth = 0.0f;
        float cc = (hinv * erp) * depth;
        if (cc > 3.0f) cc = 3.0f;                                     /* ContactMaxCorrectingVel (PhysicsEngineODE.cpp:26) */
        const V3 ra = cross(r, nrm);
        const float outgoing = dot(nrm, C.v) + dot(ra, C.w);
        const float bounce = wall ? 0.01f : 0.0f;
        if (-outgoing > 0.0f) { const float newc = -bounce * outgoing; if (newc > cc) cc = newc; }
        V3 t1, t2; plane_space(nrm, t1, t2);
        const V3 d3[3] = {nrm, t1, t2};
        for (int q = 0; q < 3; ++q) {
            const int k = 3 * i + q;
            const V3 a = cross(r, d3[q]);
            J[k][0] = d3[q].x; J[k][1] = d3[q].y; J[k][2] = d3[q].z; J[k][3] = a.x; J[k][4] = a.y; J[k][5] = a.z;
            float s = 0.0f;
            for (int u = 0; u < 6; ++u) s += J[k][u] * dC.t1[u];
            rc[k] = (q == 0 ? cc : 0.0f) * hinv - s;
            cfm[k] = q == 0 ? (wall ? 0.0001f : 0.000952380942f) : P.worldCFM;
        }
        mu[i] = wall ? 0.25f : 0.1f;
        if (fresh && wall) {      /* Car::onCollisionCallback (Car.cpp:960-999); floor contacts meet group 1 (TRACK): no damage */
            const V3 vp = C.v + cross(C.w, r);
            const float rel = -(dot(vp, nrm) * 3.6f);
            const float fDamage = rel * P.mechanicalDamageRate;
            if (rel > 0.0f) {
                if (rel * P.mechanicalDamageRate > 150.0f) lifeLeft = -100.0f;            /* Engine::blowUp */
                const V3 vn = norm(to_local(C.fr, pos));
                const V3 pl = to_local(C.fr, pos);
                int zone;
                if (fabsf(vn.z) <= 0.70700002f) zone = (pl.x >= 0.0f) ? 2 : 3; else zone = (pl.z <= 0.0f) ? 1 : 0;
                dmg[zone] = tmaxf(dmg[zone], fDamage); dmg[4] = tmaxf(dmg[4], fDamage);
            }
        }
    }
    for (int k = 0; k < m; ++k) {
        solve6(S21, J[k], W[k]);                                      /* W_k = S^-1 J_k^T */
        float s = 0.0f;
        for (int u = 0; u < 6; ++u) s += J[k][u] * z0[u];
        rc[k] -= s;
        lam[k] = 0.0f;
    }
    for (int i = 0; i < m; ++i)
        for (int k = 0; k < m; ++k) {
            float s = 0.0f;
            for (int u = 0; u < 6; ++u) s += J[i][u] * W[k][u];
            A[i][k] = s + (i == k ? cfm[i] * hinv : 0.0f);
        }
    for (int it = 0; it < 50; ++it)
        for (int i = 0; i < m; ++i) {
            float s = rc[i];
            for (int k = 0; k < m; ++k) s -= A[i][k] * lam[k];
            float v = lam[i] + s / A[i][i];
            float lo = 0.0f, hi = 3.4e38f;
            if (i % 3 != 0) { hi = mu[i / 3] * lam[i - i % 3]; lo = -hi; }
            if (v < lo) v = lo;
            if (v > hi) v = hi;
            lam[i] = v;
        }
    for (int u = 0; u < 6; ++u) { float s = z0[u]; for (int k = 0; k < m; ++k) s += W[k][u] * lam[k]; z[u] = s; }
}

} // namespace pd
