/*
 * oracle/ode_restate/ode_collide.h -- TEST INFRASTRUCTURE (oracle only).
 *
 * Contact DETECTION between the car's colliders and the static track meshes, standing in for the ODE 0.16.3
 * colliders the reference reaches through dCollide (Physics/ODE/PhysicsEngineODE.cpp:246-282):
 *   box  vs trimesh  (car floor box  vs TRACK meshes)  ode/src/collision_trimesh_box.cpp   (dCollideBTL)
 *   mesh vs trimesh  (car hull mesh  vs WALL  meshes)  ode/src/collision_trimesh_trimesh.cpp (dCollideTTL)
 * ODE's source is absent from /root/reference, and the reference holds no test vector for this path:
 * PARITY UNPINNED.  What is restated is what the reference's own code consumes:
 *   - whether a (geom, geom) pair yields at least one contact (PhysicsEngineODE.cpp:276 `n > 0`), and
 *   - for box-vs-trimesh, the contact normal, which PhysicsEngineODE::onCollision filters by its chassis-local
 *     y component (>= 0.9 kept, :309-318).
 * box vs triangle: separating-axis test over the 13 axes of collision_trimesh_box.cpp (_cldTestSeparatingAxes):
 * the triangle normal (one-sided: depth = r + (v0 - c).n), the 3 box axes and the 9 edge cross products, the axis of
 * least depth wins, edge axes only when 1.5 x depth is still smaller (the bias of _cldTestEdge); the reported normal
 * points from the triangle towards the box.  mesh vs mesh: a pair of triangles touches when an edge of either crosses
 * the other (the edge / triangle clipping both ODE trimesh-trimesh colliders are built on); coplanar overlap is not
 * detected.
 *
 * Contact GENERATION (positions, normals, depths) -- ODE's own generators (the clipping of collision_trimesh_box.cpp, the
 * OPCODE pair walk of collision_trimesh_trimesh.cpp) cannot be restated from memory to the vertex, and nothing in the reference
 * pins them, so the contacts handed to the contact joints are defined HERE, as simply as the physics allows, and the product
 * follows this definition (pd_contacts.h):
 *   floor box vs an accepted TRACK triangle (SAT overlap + body-local normal filter): every box corner that lies BELOW the
 *     triangle's plane and whose projection falls inside the triangle is a contact (position = the corner, normal = the
 *     triangle's unit normal towards the box, depth = distance below the plane); per corner the deepest triangle wins; when no
 *     corner qualifies for any accepted triangle, one contact at the box's support corner along the SAT normal of the deepest
 *     accepted triangle with the SAT depth.
 *   hull mesh vs a WALL triangle that at least one hull triangle crosses: every hull VERTEX that lies behind the wall triangle's
 *     plane (seen from the chassis origin) by less than 0.5 m and projects inside it is a contact (position = the vertex, normal
 *     = the wall's unit normal towards the chassis, depth = distance behind the plane); per vertex the deepest triangle wins.
 *   At most 4 floor and 4 wall contacts per car, the deepest first (ties: lower corner / vertex index).
 */
#pragma once
#include <cmath>
#include <vector>

namespace oder {

struct CV3 { float x, y, z; };
static inline CV3 cv(float x, float y, float z) { CV3 r = {x, y, z}; return r; }
static inline CV3 csub(CV3 a, CV3 b) { return cv(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float cdot(CV3 a, CV3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline CV3 ccross(CV3 a, CV3 b) { return cv(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

/* one candidate axis L (not normalised): p[k] = (v_k - c).L, box radius r; updates (best depth, best normal) */
static inline bool sat_axis(CV3 L, const float p[3], float r, float bias, float& bestDepth, CV3& bestN) {
    const float fMin = fminf(p[0], fminf(p[1], p[2])), fMax = fmaxf(p[0], fmaxf(p[1], p[2]));
    if (fMin > r || fMax < -r) return false;                 /* separated */
    const float len = sqrtf(cdot(L, L));
    if (!(len > 1e-6f)) return true;                         /* degenerate axis: cannot separate, cannot be the normal */
    const float dMin = r - fMin, dMax = fMax + r;
    float depth; float sgn;
    if (dMin > dMax) { depth = dMax; sgn = 1.0f; } else { depth = dMin; sgn = -1.0f; }
    const float inv = 1.0f / len;
    depth *= inv;
    if (depth * bias < bestDepth) { bestDepth = depth; bestN = cv(L.x * inv * sgn, L.y * inv * sgn, L.z * inv * sgn); }
    return true;
}

/* box: centre c, world axes A[3] (unit), half sizes h[3]; triangle v0 v1 v2.  Returns true when they overlap and
 * writes the contact normal (unit, pointing from the triangle towards the box). */
static inline bool box_tri_contact(CV3 c, const CV3 A[3], const float h[3], CV3 v0, CV3 v1, CV3 v2, CV3& nOut, float* depthOut = nullptr) {
    const CV3 E[3] = {csub(v1, v0), csub(v2, v1), csub(v0, v2)};
    const CV3 P[3] = {csub(v0, c), csub(v1, c), csub(v2, c)};
    const CV3 N = ccross(E[0], csub(v2, v0));
    float bestDepth = 3.4e38f; CV3 bestN = cv(0, 0, 0);
    { /* axis 1: the triangle's normal, one-sided */
        const float len = sqrtf(cdot(N, N));
        if (!(len > 1e-12f)) return false;                   /* degenerate triangle */
        const float r = h[0] * fabsf(cdot(A[0], N)) + h[1] * fabsf(cdot(A[1], N)) + h[2] * fabsf(cdot(A[2], N));
        const float depth = r + cdot(P[0], N);
        if (depth < 0.0f) return false;
        const float inv = 1.0f / len;
        bestDepth = depth * inv; bestN = cv(N.x * inv, N.y * inv, N.z * inv);
    }
    for (int i = 0; i < 3; ++i) { /* axes 2-4: the box's faces */
        const float p[3] = {cdot(P[0], A[i]), cdot(P[1], A[i]), cdot(P[2], A[i])};
        if (!sat_axis(A[i], p, h[i], 1.0f, bestDepth, bestN)) return false;
    }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { /* axes 5-13: box axis x triangle edge */
        const CV3 L = ccross(A[i], E[j]);
        const float p[3] = {cdot(P[0], L), cdot(P[1], L), cdot(P[2], L)};
        const float r = h[0] * fabsf(cdot(A[0], L)) + h[1] * fabsf(cdot(A[1], L)) + h[2] * fabsf(cdot(A[2], L));
        if (!sat_axis(L, p, r, 1.5f, bestDepth, bestN)) return false;
    }
    nOut = bestN;
    if (depthOut) *depthOut = bestDepth;
    return true;
}

/* is the projection of p onto the plane of (v0, v1, v2) along its normal inside the triangle?  (edge functions on the
 * UNNORMALISED normal N = (v1 - v0) x (v2 - v0); boundary counts as inside) */
static inline bool projects_inside(CV3 p, CV3 v0, CV3 v1, CV3 v2, CV3 N) {
    const float d0 = cdot(ccross(csub(v1, v0), csub(p, v0)), N);
    const float d1 = cdot(ccross(csub(v2, v1), csub(p, v1)), N);
    const float d2 = cdot(ccross(csub(v0, v2), csub(p, v2)), N);
    return d0 >= 0.0f && d1 >= 0.0f && d2 >= 0.0f;
}

struct ContactPoint { CV3 pos, normal; float depth; int kind; };     /* kind 0: floor box vs TRACK (mode 28700), 1: hull vs WALL (mode 28692) */

/* candidate for slot `key` (box corner / hull vertex): deeper wins; ties by larger normal.y, then .x, then .z */
static inline void contact_offer(ContactPoint* slot, bool* used, int key, CV3 pos, CV3 n, float depth, int kind) {
    ContactPoint& c = slot[key];
    bool take = !used[key];
    if (!take) {
        if (depth > c.depth) take = true;
        else if (depth == c.depth) {
            if (n.y > c.normal.y) take = true;
            else if (n.y == c.normal.y && (n.x > c.normal.x || (n.x == c.normal.x && n.z > c.normal.z))) take = true;
        }
    }
    if (take) { c.pos = pos; c.normal = n; c.depth = depth; c.kind = kind; used[key] = true; }
}
/* the `maxOut` deepest used slots, deepest first, ties by lower key */
static inline int contact_select(const ContactPoint* slot, const bool* used, int nSlots, ContactPoint* out, int maxOut) {
    int n = 0; std::vector<char> taken((size_t)(nSlots > 0 ? nSlots : 1), 0);
    for (int r = 0; r < maxOut; ++r) {
        int best = -1;
        for (int k = 0; k < nSlots; ++k) if (used[k] && !taken[k] && (best < 0 || slot[k].depth > slot[best].depth)) best = k;
        if (best < 0) break;
        taken[best] = true; out[n++] = slot[best];
    }
    return n;
}

/* segment p -> q against triangle (v0, e1, e2), both faces */
static inline bool seg_tri(CV3 p, CV3 q, CV3 v0, CV3 e1, CV3 e2) {
    const CV3 d = csub(q, p);
    const CV3 pvec = ccross(d, e2);
    const float det = cdot(e1, pvec);
    if (fabsf(det) < 1e-12f) return false;
    const float inv = 1.0f / det;
    const CV3 tvec = csub(p, v0);
    const float u = cdot(tvec, pvec) * inv;
    if (u < 0.0f || u > 1.0f) return false;
    const CV3 qvec = ccross(tvec, e1);
    const float v = cdot(d, qvec) * inv;
    if (v < 0.0f || u + v > 1.0f) return false;
    const float t = cdot(e2, qvec) * inv;
    return t >= 0.0f && t <= 1.0f;
}

static inline bool tri_tri(CV3 a0, CV3 a1, CV3 a2, CV3 b0, CV3 b1, CV3 b2) {
    const CV3 ae1 = csub(a1, a0), ae2 = csub(a2, a0), be1 = csub(b1, b0), be2 = csub(b2, b0);
    return seg_tri(a0, a1, b0, be1, be2) || seg_tri(a1, a2, b0, be1, be2) || seg_tri(a2, a0, b0, be1, be2) ||
           seg_tri(b0, b1, a0, ae1, ae2) || seg_tri(b1, b2, a0, ae1, ae2) || seg_tri(b2, b0, a0, ae1, ae2);
}

} // namespace oder
