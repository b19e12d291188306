/*
 * pd_collide.h -- contact DETECTION between the chassis colliders and the static track meshes:
 * PhysicsEngineODE::collisionStep / collisionNearCallback / onCollision (Physics/ODE/PhysicsEngineODE.cpp:228-341)
 * as far as Car::onCollisionCallback (Car/Car.cpp:921-1044) turns it into `collisionFlag`.
 *
 *   floor box  (colliders.ini, category CAR, mask TRACK)  vs  TRACK triangles: 13-axis separating-axis test
 *       (triangle normal one-sided, 3 box axes, 9 edge cross products; least depth wins, edge axes with the 1.5 bias);
 *       a contact counts only when its normal, in chassis coordinates, has y >= 0.9 (PhysicsEngineODE.cpp:309-318)
 *   hull mesh  (collider.bin,   category CAR, mask WALL)   vs  WALL triangles: a triangle pair touches when an edge
 *       of either crosses the other
 * Tested on odd physics frames only (PhysicsEngineODE.cpp:230-236; even frames are dynamic-vs-dynamic, and there is
 * one car per env).  The response (contact generation + contact joints) lives in pd_contacts.h.
 *
 * Hull vs wall is evaluated in the chassis frame (the hull's model space, as OPCODE's mesh-vs-mesh query does): the
 * hull keeps its chassis-local vertices, each candidate wall triangle is transformed into that frame.
 *
 * Broad phase: a uniform x-z grid (2 m cells) with two triangle lists per cell (TRACK, WALL) and the height range of
 * each list -- on the open road a cell costs one 16-byte load; then the triangle's box against the collider's world
 * box; for hull-vs-wall the transformed triangle's box against the hull's bounds and against each hull triangle's
 * bounds (precomputed, PdCarParams::colliderTriBounds).  Every stage only removes pairs that cannot touch, so the
 * flag equals the all-pairs answer.
 */
#pragma once
#include "pd_track.h"

namespace pd {

/* one separating-axis candidate L (not normalised): p = projections of the triangle's vertices (relative to the box
 * centre), r = the box' radius along L.  false = separated. */
PD_HD bool sat_axis(V3 L, float p0, float p1, float p2, float r, float bias, float& bestDepth, V3& bestN) {
    const float fMin = fminf(p0, fminf(p1, p2)), fMax = fmaxf(p0, fmaxf(p1, p2));
    if (fMin > r || fMax < -r) return false;
    const float len = sqrtf(dot(L, L));
    if (!(len > 1e-6f)) return true;
    const float dMin = r - fMin, dMax = fMax + r;
    float depth, sgn;
    if (dMin > dMax) { depth = dMax; sgn = 1.0f; } else { depth = dMin; sgn = -1.0f; }
    const float inv = 1.0f / len;
    depth *= inv;
    if (depth * bias < bestDepth) { bestDepth = depth; bestN = v3(L.x * inv * sgn, L.y * inv * sgn, L.z * inv * sgn); }
    return true;
}

/* box (centre c, unit axes A0 A1 A2, half sizes h) against triangle v0 v1 v2; nOut = contact normal (triangle -> box) */
PD_HDN bool box_tri_contact(V3 c, V3 A0, V3 A1, V3 A2, V3 h, V3 v0, V3 v1, V3 v2, V3& nOut, float* depthOut = nullptr) {
    const V3 E0 = v1 - v0, E1 = v2 - v1, E2 = v0 - v2;
    const V3 P0 = v0 - c, P1 = v1 - c, P2 = v2 - c;
    const V3 N = cross(E0, v2 - v0);
    float bestDepth = 3.4e38f; V3 bestN = v3(0, 0, 0);
    {
        const float len = sqrtf(dot(N, N));
        if (!(len > 1e-12f)) return false;
        const float r = h.x * fabsf(dot(A0, N)) + h.y * fabsf(dot(A1, N)) + h.z * fabsf(dot(A2, N));
        const float depth = r + dot(P0, N);
        if (depth < 0.0f) return false;
        const float inv = 1.0f / len;
        bestDepth = depth * inv; bestN = v3(N.x * inv, N.y * inv, N.z * inv);
    }
    if (!sat_axis(A0, dot(P0, A0), dot(P1, A0), dot(P2, A0), h.x, 1.0f, bestDepth, bestN)) return false;
    if (!sat_axis(A1, dot(P0, A1), dot(P1, A1), dot(P2, A1), h.y, 1.0f, bestDepth, bestN)) return false;
    if (!sat_axis(A2, dot(P0, A2), dot(P1, A2), dot(P2, A2), h.z, 1.0f, bestDepth, bestN)) return false;
    const V3 A[3] = {A0, A1, A2}; const V3 E[3] = {E0, E1, E2};
    PD_NOUNROLL
    for (int i = 0; i < 3; ++i) {
        PD_NOUNROLL
        for (int j = 0; j < 3; ++j) {
            const V3 L = cross(A[i], E[j]);
            const float r = h.x * fabsf(dot(A0, L)) + h.y * fabsf(dot(A1, L)) + h.z * fabsf(dot(A2, L));
            if (!sat_axis(L, dot(P0, L), dot(P1, L), dot(P2, L), r, 1.5f, bestDepth, bestN)) return false;
        }
    }
    nOut = bestN;
    if (depthOut) *depthOut = bestDepth;
    return true;
}

/* segment p -> q against triangle (v0, e1, e2), both faces */
PD_HD bool seg_tri(V3 p, V3 q, V3 v0, V3 e1, V3 e2) {
    const V3 d = q - p;
    const V3 pvec = cross(d, e2);
    const float det = dot(e1, pvec);
    if (fabsf(det) < 1e-12f) return false;
    const float inv = 1.0f / det;
    const V3 tvec = p - v0;
    const float u = dot(tvec, pvec) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const V3 qvec = cross(tvec, e1);
    const float v = dot(d, qvec) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = dot(e2, qvec) * inv;
    return t >= 0.0f && t <= 1.0f;
}
/* filter: true when triangle b lies wholly on one side of triangle a's plane, by more than 0.1 mm -- then no edge of
 * either can cross the other */
PD_HD bool tri_plane_separates(V3 a0, V3 a1, V3 a2, V3 b0, V3 b1, V3 b2) {
    const V3 n = cross(a1 - a0, a2 - a0);
    const float d = dot(n, a0), margin = 1e-4f * sqrtf(dot(n, n));
    const float s0 = dot(n, b0) - d, s1 = dot(n, b1) - d, s2 = dot(n, b2) - d;
    return (s0 > margin && s1 > margin && s2 > margin) || (s0 < -margin && s1 < -margin && s2 < -margin);
}
PD_HDN bool tri_tri(V3 a0, V3 a1, V3 a2, V3 b0, V3 b1, V3 b2) {
    const V3 ae1 = a1 - a0, ae2 = a2 - a0, be1 = b1 - b0, be2 = b2 - b0;
    return seg_tri(a0, a1, b0, be1, be2) || seg_tri(a1, a2, b0, be1, be2) || seg_tri(a2, a0, b0, be1, be2) ||
           seg_tri(b0, b1, a0, ae1, ae2) || seg_tri(b1, b2, a0, ae1, ae2) || seg_tri(b2, b0, a0, ae1, ae2);
}

/* one entry of the collision grid: the triangle's box and its index */
PD_HD void load_coll_rec(const float* rec, int k, float* mn, float* mx, int& tri) {
#if defined(__CUDA_ARCH__)
    const float4 a = __ldg(reinterpret_cast<const float4*>(rec) + 2 * (size_t)k), b = __ldg(reinterpret_cast<const float4*>(rec) + 2 * (size_t)k + 1);
    mn[0] = a.x; mn[1] = a.y; mn[2] = a.z; tri = __float_as_int(a.w); mx[0] = b.x; mx[1] = b.y; mx[2] = b.z;
#else
    const float* p = rec + (size_t)k * 8;
    mn[0] = p[0]; mn[1] = p[1]; mn[2] = p[2]; memcpy(&tri, &p[3], 4); mx[0] = p[4]; mx[1] = p[5]; mx[2] = p[6];
#endif
}

/* triangle (b0 b1 b2) against the axis-aligned box [cmin, cmax], both in the same frame: the separating-axis test over the
 * triangle's normal and the 9 edge x box-axis products (the 3 box axes are the caller's min / max comparison).  A filter:
 * axes are compared with a 1e-4 slack, `true` = certainly apart. */
PD_HD bool tri_outside_aabb(V3 b0, V3 b1, V3 b2, V3 cmin, V3 cmax) {
    const V3 c = v3(0.5f * (cmin.x + cmax.x), 0.5f * (cmin.y + cmax.y), 0.5f * (cmin.z + cmax.z));
    const V3 h = v3(0.5f * (cmax.x - cmin.x) + 1e-4f, 0.5f * (cmax.y - cmin.y) + 1e-4f, 0.5f * (cmax.z - cmin.z) + 1e-4f);
    const V3 p0 = b0 - c, p1 = b1 - c, p2 = b2 - c;
    const V3 e0 = p1 - p0, e1 = p2 - p1, e2 = p0 - p2;
    {
        const V3 n = cross(e0, p2 - p0);
        const float r = h.x * fabsf(n.x) + h.y * fabsf(n.y) + h.z * fabsf(n.z);
        if (fabsf(dot(n, p0)) > r * 1.0001f + 1e-6f) return true;
    }
    const V3 E[3] = {e0, e1, e2};
    PD_UNROLL
    for (int j = 0; j < 3; ++j) {
        const V3 e = E[j];
        { /* x axis: L = (0, -e.z, e.y) */
            const float a = -e.z * p0.y + e.y * p0.z, b = -e.z * p1.y + e.y * p1.z, cc = -e.z * p2.y + e.y * p2.z;
            const float r = h.y * fabsf(e.z) + h.z * fabsf(e.y);
            if (fminf(a, fminf(b, cc)) > r * 1.0001f + 1e-6f || fmaxf(a, fmaxf(b, cc)) < -(r * 1.0001f + 1e-6f)) return true;
        }
        { /* y axis: L = (e.z, 0, -e.x) */
            const float a = e.z * p0.x - e.x * p0.z, b = e.z * p1.x - e.x * p1.z, cc = e.z * p2.x - e.x * p2.z;
            const float r = h.x * fabsf(e.z) + h.z * fabsf(e.x);
            if (fminf(a, fminf(b, cc)) > r * 1.0001f + 1e-6f || fmaxf(a, fmaxf(b, cc)) < -(r * 1.0001f + 1e-6f)) return true;
        }
        { /* z axis: L = (-e.y, e.x, 0) */
            const float a = -e.y * p0.x + e.x * p0.y, b = -e.y * p1.x + e.x * p1.y, cc = -e.y * p2.x + e.x * p2.y;
            const float r = h.x * fabsf(e.y) + h.y * fabsf(e.x);
            if (fminf(a, fminf(b, cc)) > r * 1.0001f + 1e-6f || fmaxf(a, fmaxf(b, cc)) < -(r * 1.0001f + 1e-6f)) return true;
        }
    }
    return false;
}

/* four entries k, k + step, k + 2 step, k + 3 step of a list ending at kEnd; entries past the end read as "below everything" */
struct CollRec4 { float mn[4][3], mx[4][3]; int tri[4]; };
PD_HD void load_coll_rec4(const float* rec, int k, int step, int kEnd, CollRec4& R) {
    PD_UNROLL
    for (int u = 0; u < 4; ++u) {
        const int kk = k + u * step;
        if (kk < kEnd) load_coll_rec(rec, kk, R.mn[u], R.mx[u], R.tri[u]);
        else { R.mn[u][0] = R.mn[u][1] = R.mn[u][2] = 0.0f; R.mx[u][0] = R.mx[u][2] = 0.0f; R.mx[u][1] = -3.4e38f; R.tri[u] = 0; }
    }
}

/* conservative: true when the axis-aligned box [mn, mx] lies outside the oriented box (centre c, axes of f, half sizes h
 * grown by 1 mm) along one of the oriented box' own axes */
PD_HD bool aabb_outside_obb(const float* mn, const float* mx, V3 c, const Frame& f, V3 h) {
    const V3 m = v3(0.5f * (mn[0] + mx[0]) - c.x, 0.5f * (mn[1] + mx[1]) - c.y, 0.5f * (mn[2] + mx[2]) - c.z);
    const V3 e = v3(0.5f * (mx[0] - mn[0]) + 1e-3f, 0.5f * (mx[1] - mn[1]) + 1e-3f, 0.5f * (mx[2] - mn[2]) + 1e-3f);
    if (fabsf(dot(m, f.ay)) > h.y + (e.x * fabsf(f.ay.x) + e.y * fabsf(f.ay.y) + e.z * fabsf(f.ay.z))) return true;
    if (fabsf(dot(m, f.ax)) > h.x + (e.x * fabsf(f.ax.x) + e.y * fabsf(f.ax.y) + e.z * fabsf(f.ax.z))) return true;
    if (fabsf(dot(m, f.az)) > h.z + (e.x * fabsf(f.az.x) + e.y * fabsf(f.az.y) + e.z * fabsf(f.az.z))) return true;
    return false;
}

/* world box of an oriented box: centre c, axes of frame f, half sizes h */
PD_HD void obb_bounds(const Frame& f, V3 c, V3 h, V3& lo, V3& hi) {
    const V3 e = v3(h.x * fabsf(f.ax.x) + h.y * fabsf(f.ay.x) + h.z * fabsf(f.az.x),
                    h.x * fabsf(f.ax.y) + h.y * fabsf(f.ay.y) + h.z * fabsf(f.az.y),
                    h.x * fabsf(f.ax.z) + h.y * fabsf(f.ay.z) + h.z * fabsf(f.az.z));
    lo = c - e; hi = c + e;
}

/* The floor box entirely in FRONT of the triangle's plane (by more than 0.1 mm): box_tri_contact's first axis -- the triangle's normal, one-sided:
 * depth = r + (v0 - c) . N < 0 -- answers "no contact" for such a triangle whatever the other twelve axes say, so the entry needs neither its box
 * tests nor the triangle itself.  (The road under a car: the floor box rides a few centimetres above it, and nearly every entry that reaches the
 * separating-axis test is rejected by exactly this axis.)  The margin keeps the early answer on the safe side of box_tri_contact's own rounding. */
struct Plane4 { float x, y, z, w; };
PD_HD Plane4 load_plane(const float* __restrict__ planes, int k) {      /* issued together with the entry's box: the two loads overlap */
#if defined(__CUDA_ARCH__)
    const float4 pl = __ldg(reinterpret_cast<const float4*>(planes) + k);
    Plane4 r; r.x = pl.x; r.y = pl.y; r.z = pl.z; r.w = pl.w; return r;
#else
    const float* q = planes + (size_t)k * 4; Plane4 r; r.x = q[0]; r.y = q[1]; r.z = q[2]; r.w = q[3]; return r;
#endif
}
PD_HD bool box_clear_of_plane(const Plane4& pl, const Frame& f, V3 bc, V3 bh) {
    const V3 n = v3(pl.x, pl.y, pl.z);
    if (n.x == 0.0f && n.y == 0.0f && n.z == 0.0f) return false;
    const float r = bh.x * fabsf(dot(f.ax, n)) + bh.y * fabsf(dot(f.ay, n)) + bh.z * fabsf(dot(f.az, n));
    return r + (pl.w - dot(bc, n)) < -1e-4f;
}

/* Does the chassis touch the static world?  The work is shared by cellParts x nparts callers that OR their answers:
 * caller (cellPart, part) visits the cells number cellPart, cellPart + cellParts, ... of the car's footprint and, in each,
 * the entries part, part + nparts, ... of the cell's lists.  (1 x 1: thread-per-car kernel; 1 x 4: a quad inside the tick
 * kernel; 4 x 8: a warp per car in k_collide.)
 * what: bit 0 = floor box vs TRACK triangles, bit 1 = hull vs WALL triangles.  With the walls left out, *wallsPossible tells whether any
 * cell of the footprint holds WALL triangles in the hull's height range (k_collide: the floor test runs one thread per car, the rare car
 * near a wall then gets a whole warp for the hull test). */
PD_HDN bool car_collide(const PdCarParams& P, const TrackDev& T, const Body& C, int part, int nparts, int cellPart = 0, int cellParts = 1, int what = 3, bool* wallsPossible = nullptr) {
    const PdBoundGrid& G = T.collGrid;
    if (G.nx <= 0 || G.nz <= 0) return false;
    const Frame& f = C.fr;
    /* floor box */
    const bool hasBox = P.hasBoxCollider != 0;
    const V3 bc = to_world(f, v3(P.boxCentre[0], P.boxCentre[1], P.boxCentre[2]));
    const V3 bh = v3(P.boxSize[0] * 0.5f, P.boxSize[1] * 0.5f, P.boxSize[2] * 0.5f);
    V3 blo, bhi; obb_bounds(f, bc, bh, blo, bhi);
    /* hull mesh: oriented bounding box of its chassis-local bounds, a little inflated (it is a filter only) */
    const bool hasHull = P.nColliderTris > 0;
    const V3 hcl = v3((P.colliderMin[0] + P.colliderMax[0]) * 0.5f, (P.colliderMin[1] + P.colliderMax[1]) * 0.5f, (P.colliderMin[2] + P.colliderMax[2]) * 0.5f);
    const V3 hh = v3((P.colliderMax[0] - P.colliderMin[0]) * 0.5f + 1e-3f, (P.colliderMax[1] - P.colliderMin[1]) * 0.5f + 1e-3f, (P.colliderMax[2] - P.colliderMin[2]) * 0.5f + 1e-3f);
    const V3 hc = to_world(f, hcl);
    V3 hlo, hhi; obb_bounds(f, hc, hh, hlo, hhi);
    if (!hasBox && !hasHull) return false;
    float x0 = hasBox ? blo.x : hlo.x, x1 = hasBox ? bhi.x : hhi.x, z0 = hasBox ? blo.z : hlo.z, z1 = hasBox ? bhi.z : hhi.z;
    if (hasHull) { x0 = tminf(x0, hlo.x); x1 = tmaxf(x1, hhi.x); z0 = tminf(z0, hlo.z); z1 = tmaxf(z1, hhi.z); }
    int ix0 = (int)floorf((x0 - G.ox) * G.invCell), ix1 = (int)floorf((x1 - G.ox) * G.invCell);
    int iz0 = (int)floorf((z0 - G.oz) * G.invCell), iz1 = (int)floorf((z1 - G.oz) * G.invCell);
    if (ix0 < 0) ix0 = 0; if (iz0 < 0) iz0 = 0; if (ix1 >= G.nx) ix1 = G.nx - 1; if (iz1 >= G.nz) iz1 = G.nz - 1;
    if (ix1 - ix0 > 8) ix1 = ix0 + 8;           /* a car spans at most 3-4 cells; a non-finite pose must not walk the whole grid */
    if (iz1 - iz0 > 8) iz1 = iz0 + 8;
    const V3 cmin = v3(P.colliderMin[0] - 1e-4f, P.colliderMin[1] - 1e-4f, P.colliderMin[2] - 1e-4f), cmax = v3(P.colliderMax[0] + 1e-4f, P.colliderMax[1] + 1e-4f, P.colliderMax[2] + 1e-4f);
    const int wx = ix1 - ix0 + 1, ncell = wx * (iz1 - iz0 + 1);
    for (int ci = cellPart; ci < ncell; ci += cellParts) {
        {
            const int c = (iz0 + ci / wx) * G.nx + (ix0 + ci % wx);
#if defined(__CUDA_ARCH__)
            const float4 yr = __ldg(reinterpret_cast<const float4*>(T.collCell) + 2 * (size_t)c), kk = __ldg(reinterpret_cast<const float4*>(T.collCell) + 2 * (size_t)c + 1);
            const float tY0 = yr.x, tY1 = yr.y, wY0 = yr.z, wY1 = yr.w;
            const int kT = __float_as_int(kk.x), kW = __float_as_int(kk.y), kE = __float_as_int(kk.z);
#else
            const float* cq = T.collCell + (size_t)c * 8;
            const float tY0 = cq[0], tY1 = cq[1], wY0 = cq[2], wY1 = cq[3];
            int kT, kW, kE; memcpy(&kT, &cq[4], 4); memcpy(&kW, &cq[5], 4); memcpy(&kE, &cq[6], 4);
#endif
            if (hasBox && (what & 1) && !(tY0 > bhi.y || tY1 < blo.y)) {           /* C_CATEGORY_TRACK triangles x floor box */
                const int k0 = kT, k1 = kW;
                /* four entries of this lane's share are fetched together (the loop is bound by load latency, not by arithmetic) */
                bool stop = false;
                for (int k = k0 + part; k < k1 && !stop; k += 4 * nparts) {
                    CollRec4 R; load_coll_rec4(T.collRec, k, nparts, k1, R);
                    Plane4 PL[4];
                    PD_UNROLL
                    for (int u = 0; u < 4; ++u) PL[u] = load_plane(T.collPlane, (k + u * nparts < k1) ? k + u * nparts : k);
                    PD_UNROLL
                    for (int u = 0; u < 4; ++u) {
                        if (stop) break;
                        const float* mn = R.mn[u]; const float* mx = R.mx[u];
                        if (mx[1] < blo.y) { stop = true; break; }   /* sorted by descending top: nothing further reaches the box (also ends a short group) */
                        if (mn[1] > bhi.y || mn[0] > bhi.x || mx[0] < blo.x || mn[2] > bhi.z || mx[2] < blo.z) continue;
                        if (box_clear_of_plane(PL[u], f, bc, bh)) continue;
                        if (aabb_outside_obb(mn, mx, bc, f, bh)) continue;       /* the entry's box misses the collider along one of its axes */
                        const float* q = T.triRaw + (size_t)R.tri[u] * 9;
                        const V3 v0 = v3(q[0], q[1], q[2]), v1 = v3(q[3], q[4], q[5]), v2 = v3(q[6], q[7], q[8]);
                        V3 n;
                        if (!box_tri_contact(bc, f.ax, f.ay, f.az, bh, v0, v1, v2, n)) continue;
                        if (dot(f.ay, n) < 0.9f) continue;          /* chassis-local normal y (dBodyVectorFromWorld) */
                        return true;
                    }
                }
            }
            if (hasHull && !(what & 2) && wallsPossible && kE > kW && !(wY0 > hhi.y || wY1 < hlo.y)) *wallsPossible = true;
            if (hasHull && (what & 2) && !(wY0 > hhi.y || wY1 < hlo.y)) {          /* C_CATEGORY_WALL triangles x hull mesh */
                const int k0 = kW, k1 = kE;
                bool stop = false;
                for (int k = k0 + part; k < k1 && !stop; k += 4 * nparts) {
                  CollRec4 R; load_coll_rec4(T.collRec, k, nparts, k1, R);
                  PD_UNROLL
                  for (int u = 0; u < 4; ++u) {
                    if (stop) break;
                    const float* mn = R.mn[u]; const float* mx = R.mx[u]; const int t = R.tri[u];
                    if (mx[1] < hlo.y) { stop = true; break; }
                    if (mn[1] > hhi.y || mn[0] > hhi.x || mx[0] < hlo.x || mn[2] > hhi.z || mx[2] < hlo.z) continue;
                    if (aabb_outside_obb(mn, mx, hc, f, hh)) continue;
                    const float* q = T.triRaw + (size_t)t * 9;
                    /* into the chassis frame (the hull's model space); its box against the hull's bounds */
                    const V3 b0 = to_local(f, v3(q[0], q[1], q[2])), b1 = to_local(f, v3(q[3], q[4], q[5])), b2 = to_local(f, v3(q[6], q[7], q[8]));
                    const V3 lo = v3(fminf(b0.x, fminf(b1.x, b2.x)), fminf(b0.y, fminf(b1.y, b2.y)), fminf(b0.z, fminf(b1.z, b2.z)));
                    const V3 hi = v3(fmaxf(b0.x, fmaxf(b1.x, b2.x)), fmaxf(b0.y, fmaxf(b1.y, b2.y)), fmaxf(b0.z, fmaxf(b1.z, b2.z)));
                    if (lo.x > cmax.x || hi.x < cmin.x || lo.y > cmax.y || hi.y < cmin.y || lo.z > cmax.z || hi.z < cmin.z) continue;
                    if (tri_outside_aabb(b0, b1, b2, cmin, cmax)) continue;
                    /* the wall triangle's plane: a hull triangle wholly on one side of it (by more than 0.1 mm) cannot touch it */
                    const V3 nW = cross(b1 - b0, b2 - b0);
                    const float nlen = sqrtf(dot(nW, nW));
                    const float dW = dot(nW, b0), margin = 1e-4f * nlen;
                    PD_NOUNROLL
                    for (int j = 0; j < P.nColliderTris; ++j) {
                        {   /* bounding sphere of the hull triangle against the wall triangle's plane, then against its box */
                            const float* sp = P.colliderTriSphere[j];
                            const float sd = nW.x * sp[0] + nW.y * sp[1] + nW.z * sp[2] - dW;
                            if (fabsf(sd) > sp[3] * nlen + margin) continue;
                            if (lo.x > sp[0] + sp[3] || hi.x < sp[0] - sp[3] || lo.y > sp[1] + sp[3] || hi.y < sp[1] - sp[3] || lo.z > sp[2] + sp[3] || hi.z < sp[2] - sp[3]) continue;
                        }
                        const float* p0 = P.colliderVerts[P.colliderTris[j][0]]; const float* p1 = P.colliderVerts[P.colliderTris[j][1]]; const float* p2 = P.colliderVerts[P.colliderTris[j][2]];
                        const V3 a0 = v3(p0[0], p0[1], p0[2]), a1 = v3(p1[0], p1[1], p1[2]), a2 = v3(p2[0], p2[1], p2[2]);
                        const float s0 = dot(nW, a0) - dW, s1 = dot(nW, a1) - dW, s2 = dot(nW, a2) - dW;
                        if ((s0 > margin && s1 > margin && s2 > margin) || (s0 < -margin && s1 < -margin && s2 < -margin)) continue;
                        const float* tb = P.colliderTriBounds[j];
                        if (lo.x > tb[3] || hi.x < tb[0] || lo.y > tb[4] || hi.y < tb[1] || lo.z > tb[5] || hi.z < tb[2]) continue;
                        if (tri_plane_separates(a0, a1, a2, b0, b1, b2)) continue;
                        if (tri_tri(a0, a1, a2, b0, b1, b2)) return true;
                    }
                  }
                }
            }
        }
    }
    return false;
}

#if defined(__CUDACC__)
/* The same answer, computed by the 32 lanes of ONE WARP for one car (k_collide).  The warp walks the cells of the car's
 * footprint together (all cell headers arrive in one round trip: lane i fetches cell i), the entries of a cell's lists are
 * dealt to the 32 lanes, and every lane takes its own surviving wall triangle through the hull's triangles, whose filter
 * data (bounding spheres, boxes) and vertices are staged in the warp's shared memory the first time a survivor appears
 * (hullS; a caller without room for it passes null and the data is read from the kernel parameters). */
#define PD_HULLS_SPHERE 0
#define PD_HULLS_BOUNDS (PD_HULLS_SPHERE + PD_MAX_COLLIDER_TRIS * 4)
#define PD_HULLS_TRIS   (PD_HULLS_BOUNDS + PD_MAX_COLLIDER_TRIS * 6)
#define PD_HULLS_VERTS  (PD_HULLS_TRIS + PD_MAX_COLLIDER_TRIS)
#define PD_HULLS_WORDS  (PD_HULLS_VERTS + PD_MAX_COLLIDER_VERTS * 3)      /* 2496 words = 10 KB per warp */
template <bool SMEM, bool FLOOR = true> __device__ __noinline__ bool car_collide_warp(const PdCarParams& P, const TrackDev& T, const Body& C, int lane, float* hullS, int* stats = nullptr) {
    const unsigned FULL = 0xffffffffu;
    bool staged = false;
    const PdBoundGrid& G = T.collGrid;
    if (G.nx <= 0 || G.nz <= 0) return false;
    const Frame& f = C.fr;
    const bool hasBox = P.hasBoxCollider != 0, hasHull = P.nColliderTris > 0;
    if (!hasBox && !hasHull) return false;
    const V3 bc = to_world(f, v3(P.boxCentre[0], P.boxCentre[1], P.boxCentre[2]));
    const V3 bh = v3(P.boxSize[0] * 0.5f, P.boxSize[1] * 0.5f, P.boxSize[2] * 0.5f);
    V3 blo, bhi; obb_bounds(f, bc, bh, blo, bhi);
    const V3 hcl = v3((P.colliderMin[0] + P.colliderMax[0]) * 0.5f, (P.colliderMin[1] + P.colliderMax[1]) * 0.5f, (P.colliderMin[2] + P.colliderMax[2]) * 0.5f);
    const V3 hh = v3((P.colliderMax[0] - P.colliderMin[0]) * 0.5f + 1e-3f, (P.colliderMax[1] - P.colliderMin[1]) * 0.5f + 1e-3f, (P.colliderMax[2] - P.colliderMin[2]) * 0.5f + 1e-3f);
    const V3 hc = to_world(f, hcl);
    V3 hlo, hhi; obb_bounds(f, hc, hh, hlo, hhi);
    const V3 cmin = v3(P.colliderMin[0] - 1e-4f, P.colliderMin[1] - 1e-4f, P.colliderMin[2] - 1e-4f), cmax = v3(P.colliderMax[0] + 1e-4f, P.colliderMax[1] + 1e-4f, P.colliderMax[2] + 1e-4f);
    float x0 = hasBox ? blo.x : hlo.x, x1 = hasBox ? bhi.x : hhi.x, z0 = hasBox ? blo.z : hlo.z, z1 = hasBox ? bhi.z : hhi.z;
    if (hasHull) { x0 = tminf(x0, hlo.x); x1 = tmaxf(x1, hhi.x); z0 = tminf(z0, hlo.z); z1 = tmaxf(z1, hhi.z); }
    int ix0 = (int)floorf((x0 - G.ox) * G.invCell), ix1 = (int)floorf((x1 - G.ox) * G.invCell);
    int iz0 = (int)floorf((z0 - G.oz) * G.invCell), iz1 = (int)floorf((z1 - G.oz) * G.invCell);
    if (ix0 < 0) ix0 = 0; if (iz0 < 0) iz0 = 0; if (ix1 >= G.nx) ix1 = G.nx - 1; if (iz1 >= G.nz) iz1 = G.nz - 1;
    if (ix1 - ix0 > 8) ix1 = ix0 + 8;
    if (iz1 - iz0 > 8) iz1 = iz0 + 8;
    const int wx = ix1 - ix0 + 1, ncell = wx * (iz1 - iz0 + 1);
    if (stats && lane == 0) stats[0] = ncell;
    /* parked survivors of the wall lists (one per lane, chassis frame) and their narrow phase */
    bool pend = false; V3 pb0 = v3(0, 0, 0), pb1 = pb0, pb2 = pb0;
    auto narrow = [&]() -> bool {
        const long long tc0 = stats ? clock64() : 0;
                        const long long ts0 = stats ? clock64() : 0;
                        if (SMEM && !staged) {      /* first survivor of this car: the hull's filter data and vertices move to this warp's shared memory */
                            for (int i = lane; i < P.nColliderTris; i += 32) {
                                PD_UNROLL for (int q = 0; q < 4; ++q) hullS[PD_HULLS_SPHERE + i * 4 + q] = P.colliderTriSphere[i][q];
                                PD_UNROLL for (int q = 0; q < 6; ++q) hullS[PD_HULLS_BOUNDS + i * 6 + q] = P.colliderTriBounds[i][q];
                                hullS[PD_HULLS_TRIS + i] = __int_as_float((int)P.colliderTris[i][0] | ((int)P.colliderTris[i][1] << 8) | ((int)P.colliderTris[i][2] << 16));
                            }
                            for (int i = lane; i < P.nColliderVerts; i += 32) { PD_UNROLL for (int q = 0; q < 3; ++q) hullS[PD_HULLS_VERTS + i * 3 + q] = P.colliderVerts[i][q]; }
                            __syncwarp(FULL);
                            staged = true;
                            if (stats && lane == 0) stats[5] += (int)((clock64() - ts0) >> 4);
                        }
                        /* Narrow phase in two passes.  Pass 1: every lane takes its own survivor through the hull's triangles with the
                         * cheap filters only (bounding sphere vs the wall triangle's plane, box vs box) and records the hull triangles
                         * that remain as a bit mask -- the same instructions on every lane.  Pass 2: the surviving (wall triangle, hull
                         * triangle) PAIRS of the whole warp are dealt evenly to the 32 lanes, so the expensive plane / edge tests run
                         * ceil(pairs / 32) times instead of once per hull triangle any lane still cared about. */
                        unsigned pm[PD_MAX_COLLIDER_TRIS / 32];
                        PD_UNROLL for (int q = 0; q < PD_MAX_COLLIDER_TRIS / 32; ++q) pm[q] = 0u;
                        if (pend) {
                            const V3 lo = v3(fminf(pb0.x, fminf(pb1.x, pb2.x)), fminf(pb0.y, fminf(pb1.y, pb2.y)), fminf(pb0.z, fminf(pb1.z, pb2.z)));
                            const V3 hi = v3(fmaxf(pb0.x, fmaxf(pb1.x, pb2.x)), fmaxf(pb0.y, fmaxf(pb1.y, pb2.y)), fmaxf(pb0.z, fmaxf(pb1.z, pb2.z)));
                            const V3 nW = cross(pb1 - pb0, pb2 - pb0);
                            const float nlen = sqrtf(dot(nW, nW)), dW = dot(nW, pb0), margin = 1e-4f * nlen;
                            PD_NOUNROLL
                            for (int j = 0; j < P.nColliderTris; ++j) {
                                float4 sp;
                                if (SMEM) sp = *reinterpret_cast<const float4*>(hullS + PD_HULLS_SPHERE + j * 4);
                                else sp = make_float4(P.colliderTriSphere[j][0], P.colliderTriSphere[j][1], P.colliderTriSphere[j][2], P.colliderTriSphere[j][3]);   /* constant bank, uniform index */
                                const float sd = nW.x * sp.x + nW.y * sp.y + nW.z * sp.z - dW;
                                if (fabsf(sd) > sp.w * nlen + margin) continue;
                                float tb[6];
                                if (SMEM) { PD_UNROLL for (int q = 0; q < 6; ++q) tb[q] = hullS[PD_HULLS_BOUNDS + j * 6 + q]; }
                                else { PD_UNROLL for (int q = 0; q < 6; ++q) tb[q] = P.colliderTriBounds[j][q]; }
                                if (lo.x > tb[3] || hi.x < tb[0] || lo.y > tb[4] || hi.y < tb[1] || lo.z > tb[5] || hi.z < tb[2]) continue;
                                PD_UNROLL for (int q = 0; q < PD_MAX_COLLIDER_TRIS / 32; ++q) if ((j >> 5) == q) pm[q] |= 1u << (j & 31);
                            }
                        }
                        int cnt = 0;
                        PD_UNROLL for (int q = 0; q < PD_MAX_COLLIDER_TRIS / 32; ++q) cnt += __popc(pm[q]);
                        int incl = cnt;                                  /* inclusive prefix sum of the pair counts over the lanes */
                        PD_UNROLL for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
                        const int total = __shfl_sync(FULL, incl, 31);
                        bool hit = false;
                        for (int base = 0; base < total; base += 32) {
                            const int pidx = base + lane; const bool live = pidx < total;
                            const int pq = live ? pidx : total - 1;
                            int c = 0;                                   /* owner lane: the first one whose inclusive count exceeds pq */
                            PD_UNROLL for (int step = 16; step >= 1; step >>= 1) { const int v = __shfl_sync(FULL, incl, c + step - 1); if (v <= pq) c += step; }
                            const int k = pq - (__shfl_sync(FULL, incl, c) - __shfl_sync(FULL, cnt, c));      /* k-th surviving hull triangle of that lane */
                            int j = 0, kk = k; bool found = false;
                            PD_UNROLL
                            for (int q = 0; q < PD_MAX_COLLIDER_TRIS / 32; ++q) {
                                const unsigned mq = __shfl_sync(FULL, pm[q], c); const int pc = __popc(mq);
                                if (!found) { if (kk < pc) { j = q * 32 + (int)__fns(mq, 0, kk + 1); found = true; } else kk -= pc; }
                            }
                            const V3 w0 = v3(__shfl_sync(FULL, pb0.x, c), __shfl_sync(FULL, pb0.y, c), __shfl_sync(FULL, pb0.z, c));
                            const V3 w1 = v3(__shfl_sync(FULL, pb1.x, c), __shfl_sync(FULL, pb1.y, c), __shfl_sync(FULL, pb1.z, c));
                            const V3 w2 = v3(__shfl_sync(FULL, pb2.x, c), __shfl_sync(FULL, pb2.y, c), __shfl_sync(FULL, pb2.z, c));
                            if (live && found) {
                                const float* tbl = SMEM ? hullS : T.hullTables;      /* per-lane hull triangle: shared memory, or the read-only global copy */
                                const int tri = __float_as_int(tbl[PD_HULLS_TRIS + j]);
                                const float* p0 = tbl + PD_HULLS_VERTS + (tri & 255) * 3; const float* p1 = tbl + PD_HULLS_VERTS + ((tri >> 8) & 255) * 3; const float* p2 = tbl + PD_HULLS_VERTS + ((tri >> 16) & 255) * 3;
                                const V3 a0 = v3(p0[0], p0[1], p0[2]), a1 = v3(p1[0], p1[1], p1[2]), a2 = v3(p2[0], p2[1], p2[2]);
                                if (!tri_plane_separates(w0, w1, w2, a0, a1, a2) && !tri_plane_separates(a0, a1, a2, w0, w1, w2) && tri_tri(a0, a1, a2, w0, w1, w2)) hit = true;
                            }
                            if (__any_sync(FULL, hit)) break;
                        }
        pend = false;
        if (stats && lane == 0) stats[4] += (int)((clock64() - tc0) >> 4);
        return __any_sync(FULL, hit);
    };
    for (int cbase = 0; cbase < ncell; cbase += 32) {
        /* lane i fetches the header of cell cbase + i */
        float4 hy = make_float4(3.4e38f, -3.4e38f, 3.4e38f, -3.4e38f), hk = make_float4(0, 0, 0, 0);
        if (cbase + lane < ncell) {
            const int ci = cbase + lane, c = (iz0 + ci / wx) * G.nx + (ix0 + ci % wx);
            hy = __ldg(reinterpret_cast<const float4*>(T.collCell) + 2 * (size_t)c); hk = __ldg(reinterpret_cast<const float4*>(T.collCell) + 2 * (size_t)c + 1);
        }
        const int nhere = min(32, ncell - cbase);
        for (int i = 0; i < nhere; ++i) {
            const float tY0 = __shfl_sync(FULL, hy.x, i), tY1 = __shfl_sync(FULL, hy.y, i), wY0 = __shfl_sync(FULL, hy.z, i), wY1 = __shfl_sync(FULL, hy.w, i);
            const int kT = __float_as_int(__shfl_sync(FULL, hk.x, i)), kW = __float_as_int(__shfl_sync(FULL, hk.y, i)), kE = __float_as_int(__shfl_sync(FULL, hk.z, i));
            if (FLOOR && hasBox && !(tY0 > bhi.y || tY1 < blo.y)) {           /* TRACK triangles x floor box: one entry per lane and round */
                for (int kb = kT; kb < kW; kb += 32) {
                    if (stats && lane == 0) stats[1]++;
                    const int k = kb + lane;
                    bool hit = false, below = true;
                    if (k < kW) {
                        float mn[3], mx[3]; int t;
                        load_coll_rec(T.collRec, k, mn, mx, t);
                        const Plane4 pl = load_plane(T.collPlane, k);
                        below = mx[1] < blo.y;
                        if (!below && !(mn[1] > bhi.y || mn[0] > bhi.x || mx[0] < blo.x || mn[2] > bhi.z || mx[2] < blo.z) && !box_clear_of_plane(pl, f, bc, bh) && !aabb_outside_obb(mn, mx, bc, f, bh)) {
                            const float* q = T.triRaw + (size_t)t * 9;
                            V3 n;
                            if (box_tri_contact(bc, f.ax, f.ay, f.az, bh, v3(q[0], q[1], q[2]), v3(q[3], q[4], q[5]), v3(q[6], q[7], q[8]), n) && !(dot(f.ay, n) < 0.9f)) hit = true;
                        }
                    }
                    if (__any_sync(FULL, hit)) return true;
                    if (__all_sync(FULL, below)) break;              /* sorted by descending top */
                }
            }
            if (hasHull && !(wY0 > hhi.y || wY1 < hlo.y)) {          /* WALL triangles x hull */
                for (int kb = kW; kb < kE; kb += 32) {
                    if (stats && lane == 0) stats[2]++;
                    const int k = kb + lane;
                    bool cand = false, below = true;
                    V3 b0 = v3(0, 0, 0), b1 = b0, b2 = b0;
                    if (k < kE) {
                        float mn[3], mx[3]; int t;
                        load_coll_rec(T.collRec, k, mn, mx, t);
                        below = mx[1] < hlo.y;
                        if (!below && !(mn[1] > hhi.y || mn[0] > hhi.x || mx[0] < hlo.x || mn[2] > hhi.z || mx[2] < hlo.z) && !aabb_outside_obb(mn, mx, hc, f, hh)) {
                            const float* q = T.triRaw + (size_t)t * 9;
                            b0 = to_local(f, v3(q[0], q[1], q[2])); b1 = to_local(f, v3(q[3], q[4], q[5])); b2 = to_local(f, v3(q[6], q[7], q[8]));
                            const V3 lo = v3(fminf(b0.x, fminf(b1.x, b2.x)), fminf(b0.y, fminf(b1.y, b2.y)), fminf(b0.z, fminf(b1.z, b2.z)));
                            const V3 hi = v3(fmaxf(b0.x, fmaxf(b1.x, b2.x)), fmaxf(b0.y, fmaxf(b1.y, b2.y)), fmaxf(b0.z, fmaxf(b1.z, b2.z)));
                            cand = !(lo.x > cmax.x || hi.x < cmin.x || lo.y > cmax.y || hi.y < cmin.y || lo.z > cmax.z || hi.z < cmin.z) && !tri_outside_aabb(b0, b1, b2, cmin, cmax);
                        }
                    }
                    const unsigned candMask = __ballot_sync(FULL, cand);
                    if (stats && lane == 0) stats[3] += __popc(candMask);
                    /* survivors are parked, one per lane, until a full warp of them has gathered (or the walk ends): the narrow
                       phase then runs with every lane busy instead of once per round with a few */
                    if (candMask) {
                        const unsigned freeMask = ~__ballot_sync(FULL, pend);
                        if (__popc(candMask) > __popc(freeMask)) { if (narrow()) return true; }
                        const unsigned freeNow = ~__ballot_sync(FULL, pend);
                        const int rank = __popc(freeNow & ((1u << lane) - 1u));          /* this lane's rank among the free lanes */
                        const bool take = !pend && rank < __popc(candMask);
                        const int src = take ? (int)__fns(candMask, 0, rank + 1) : lane;
                        const V3 n0 = v3(__shfl_sync(FULL, b0.x, src), __shfl_sync(FULL, b0.y, src), __shfl_sync(FULL, b0.z, src));
                        const V3 n1 = v3(__shfl_sync(FULL, b1.x, src), __shfl_sync(FULL, b1.y, src), __shfl_sync(FULL, b1.z, src));
                        const V3 n2 = v3(__shfl_sync(FULL, b2.x, src), __shfl_sync(FULL, b2.y, src), __shfl_sync(FULL, b2.z, src));
                        if (take) { pb0 = n0; pb1 = n1; pb2 = n2; pend = true; }
                    }
                    if (__all_sync(FULL, below)) break;
                }
            }
        }
    }
    if (__any_sync(FULL, pend)) { if (narrow()) return true; }
    return false;
}
#endif

} // namespace pd
