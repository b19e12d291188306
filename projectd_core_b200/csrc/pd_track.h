/*
 * pd_track.h -- track queries of the hot path: wheel / teleport ray casts against the static track mesh
 *  and the spline-side queries of Car::postStep (row A10).
 *
 * Reference anchors: Physics/ODE/PhysicsEngineODE.cpp:168-214 (ray vs every static geom, min depth),
 * Physics/ODE/RayCasterODE.cpp:11-15 (first contact, back-face cull), Sim/Track.cpp:497-562
 * (rayCastTrackBounds), :564-607 (point id / direction at distance), :579-596 (nearest point),
 * :609-699 (distance along spline), Core/Spline3d.cpp:34-75 (find_nearest_point),
 * Core/VertexHash.h:45-104 (the 27-cell neighbour query that feeds Track::nearbyPoints).
 *
 * Data layout: all triangles of all static meshes (track + wall blobs of surfaces.bin) live in ONE
 * bounding-volume hierarchy built once on the host (host/bvh_build.cpp).  Leaves hold triangles as
 * (v0, e1 = v1-v0, e2 = v2-v0) + the index of the blob (= Surface) they came from, so the
 * Moller-Trumbore test below starts from exactly the edge vectors OPCODE computes.  The whole structure
 * (driftplayground: 112 411 triangles = 4.5 MB, nodes 1.8 MB) is read-only and far smaller than the
 * 126 MB L2, so after the first touch every ray is served from L2 / L1.
 */
#pragma once
#include "pd_state_io.h"
#include <float.h>

namespace pd {

struct BvhNode {          /* 32 bytes */
    float bmin[3]; int32_t left;    /* inner: left child (right = left + 1); leaf: first triangle */
    float bmax[3]; int32_t count;   /* 0 = inner node, > 0 = number of triangles in the leaf */
};

#ifndef PD_TRI_STRIDE
#define PD_TRI_STRIDE 12
#endif
struct Tri12 { V3 v0, e1, e2; int surf; };
PD_HD Tri12 load_tri(const float* tris, int t) {
    Tri12 r;
#if defined(__CUDA_ARCH__)
    const float4* q = reinterpret_cast<const float4*>(tris) + 3 * (size_t)t;
    const float4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
    r.v0 = v3(a.x, a.y, a.z); r.e1 = v3(a.w, b.x, b.y); r.e2 = v3(b.z, b.w, c.x); r.surf = __float_as_int(c.y);
#else
    const float* p = tris + (size_t)t * PD_TRI_STRIDE;
    r.v0 = v3(p[0], p[1], p[2]); r.e1 = v3(p[3], p[4], p[5]); r.e2 = v3(p[6], p[7], p[8]); memcpy(&r.surf, &p[9], 4);
#endif
    return r;
}

struct TrackDev {
    const BvhNode* nodes;
    const float* tris;        /* PD_TRI_STRIDE floats (48 B) per triangle, leaf order: v0, e1, e2, surface id bits, 0, 0 */
    const int32_t* triSurf;   /* surface (blob) index per triangle */
    const PdSurface* surfaces;
    const PdFatPoint* fat;
    const float* splineXYZ;   /* interpolated B-spline nodes */
    const float* splineDist;  /* cumulative length at node */
    const int32_t* segStart; const int32_t* segItems;   /* PdBoundGrid CSR: boundary segments per cell */
    const int32_t* ptStart; const int32_t* ptItems;     /* PdBoundGrid CSR: spline points per cell */
    const float* segRec;      /* per segItems entry, 32 B: ax, az, bx, bz, owner's best.xyz, 0 */
    const float* ptRec;       /* per ptItems entry, 16 B: best.xyz, id bits */
    PdBoundGrid grid;         /* cells of ptStart */
    PdBoundGrid segGrid;      /* cells of segStart (finer) */
    const int32_t* colStart; const int32_t* colItems;   /* vertical-ray index: triangles per x-z cell */
    PdBoundGrid colGrid;
    const float* triRaw;      /* 9 floats per triangle: v0, v1, v2 (collision detection, pd_collide.h) */
    const int32_t* collStart; const int32_t* collItems;  /* collision grid CSR, two lists per cell: TRACK triangles, WALL triangles */
    const float* collRec;     /* per entry, 32 B: box min xyz, triangle index bits | box max xyz, 0 (lists sorted by descending ymax) */
    const float* collPlane;   /* per entry, 16 B: the triangle's unit normal and plane offset (zeros: degenerate triangle, no early answer) */
    const float* collCell;    /* per cell, 32 B: track y min / max, wall y min / max | first TRACK entry, first WALL entry, end (int bits), 0 */
    PdBoundGrid collGrid;
    const float* hullTables;  /* the car hull's triangle / vertex tables in car_collide_warp's layout (PD_HULLS_*): read-only global copy */
    PdTrackInfo info;
};

struct RayHit {
    int hit;        /* hasContact */
    V3 pos, normal;
    int surface;
};

/* Closest front-facing triangle along (o, d) within `length`.
 * Triangle test = OPCODE RayTriOverlap with culling (det >= 1e-6, 0 <= u <= det, v >= 0, u+v <= det,
 * t >= 0, t < maxDist); contact = ODE dCollideRTL after the ray/trimesh swap: pos = o + d*t,
 * normal = normalize(e1 x e2). */
PD_HDN RayHit ray_cast(const TrackDev& T, V3 o, V3 d, float length) {
    RayHit h; h.hit = 0; h.surface = -1; h.pos = v3(0, 0, 0); h.normal = v3(0, 0, 0);
    if (T.info.nNodes <= 0) return h;
    float best = -1.0f; V3 bestN = v3(0, 0, 0); int bestS = -1;
    const V3 inv = v3(d.x != 0.0f ? 1.0f / d.x : 0.0f, d.y != 0.0f ? 1.0f / d.y : 0.0f, d.z != 0.0f ? 1.0f / d.z : 0.0f);
    int stack[48]; int sp = 0; stack[sp++] = 0;
    while (sp > 0) {
        const BvhNode nd = T.nodes[stack[--sp]];
        /* slab test against [0, min(length, best)] */
        float t0 = 0.0f, t1 = (best >= 0.0f) ? best : length;
        bool miss = false;
        if (d.x != 0.0f) { float a = (nd.bmin[0] - o.x) * inv.x, b = (nd.bmax[0] - o.x) * inv.x; t0 = tmaxf(t0, tminf(a, b)); t1 = tminf(t1, tmaxf(a, b)); } else if (o.x < nd.bmin[0] || o.x > nd.bmax[0]) miss = true;
        if (d.y != 0.0f) { float a = (nd.bmin[1] - o.y) * inv.y, b = (nd.bmax[1] - o.y) * inv.y; t0 = tmaxf(t0, tminf(a, b)); t1 = tminf(t1, tmaxf(a, b)); } else if (o.y < nd.bmin[1] || o.y > nd.bmax[1]) miss = true;
        if (d.z != 0.0f) { float a = (nd.bmin[2] - o.z) * inv.z, b = (nd.bmax[2] - o.z) * inv.z; t0 = tmaxf(t0, tminf(a, b)); t1 = tminf(t1, tmaxf(a, b)); } else if (o.z < nd.bmin[2] || o.z > nd.bmax[2]) miss = true;
        if (miss || t0 > t1) continue;
        if (nd.count == 0) {
            if (sp + 2 <= 48) { stack[sp++] = nd.left; stack[sp++] = nd.left + 1; }      /* always true: the builder's trees are balanced (depth <= 46 checked at load, track_loader.cpp build_bvh), the stack holds <= depth + 1 */
            continue;
        }
        for (int k = 0; k < nd.count; ++k) {
            const int t = nd.left + k;
            const Tri12 tr = load_tri(T.tris, t);
            const V3 v0 = tr.v0, e1 = tr.e1, e2 = tr.e2;
            const V3 pvec = cross(d, e2);
            const float det = dot(e1, pvec);
            if (det < 0.000001f) continue;
            const V3 tvec = o - v0;
            const float u = dot(tvec, pvec);
            if (u < 0.0f || u > det) continue;
            const V3 qvec = cross(tvec, e1);
            const float v = dot(d, qvec);
            if (v < 0.0f || u + v > det) continue;
            float dist = dot(e2, qvec);
            if (dist < 0.0f) continue;
            dist *= (1.0f / det);
            if (!(dist < length)) continue;
            if (best < 0.0f || dist < best) { best = dist; bestN = cross(e1, e2); bestS = tr.surf; }
        }
    }
    if (best >= 0.0f) {
        h.hit = 1; h.pos = v3(o.x + d.x * best, o.y + d.y * best, o.z + d.z * best);
        h.normal = norm(bestN); h.surface = bestS;
    }
    return h;
}

/* The same query for a ray pointing straight down, d = (0,-1,0) -- every ray of the hot path (wheel rays,
 * teleport ray).  One cell lookup in the column grid replaces the tree walk.  The triangle test is the test
 * above with the terms that multiply d's zero components dropped: x*0 contributes an exact (signed) zero to
 * each sum, so det, u, v and t round exactly as in ray_cast and the two functions return identical bits. */
PD_HDN RayHit ray_cast_down(const TrackDev& T, V3 o, float length) {
    RayHit h; h.hit = 0; h.surface = -1; h.pos = v3(0, 0, 0); h.normal = v3(0, 0, 0);
    const PdBoundGrid& G = T.colGrid;
    const int ix = (int)floorf((o.x - G.ox) * G.invCell), iz = (int)floorf((o.z - G.oz) * G.invCell);
    if (ix < 0 || iz < 0 || ix >= G.nx || iz >= G.nz) return h;
    const int c = iz * G.nx + ix;
    float best = -1.0f; V3 bestN = v3(0, 0, 0); int bestS = -1;
    const int k0 = T.colStart[c], k1 = T.colStart[c + 1];
    /* two-deep software pipeline: triangle indices are fetched two entries ahead, triangle data one ahead */
    int tB = (k0 + 1 < k1) ? T.colItems[k0 + 1] : 0;
    Tri12 cur; if (k0 < k1) cur = load_tri(T.tris, T.colItems[k0]);
    for (int k = k0; k < k1; ++k) {
        const int tC = (k + 2 < k1) ? T.colItems[k + 2] : 0;
        Tri12 nxt = cur; if (k + 1 < k1) nxt = load_tri(T.tris, tB);
        const V3 v0 = cur.v0, e1 = cur.e1, e2 = cur.e2; const int surf = cur.surf;
        cur = nxt; tB = tC;
        /* pvec = d x e2 = (-e2.z, 0, e2.x) */
        const float det = e1.x * (-e2.z) + e1.z * e2.x;
        if (det < 0.000001f) continue;
        const V3 tvec = o - v0;
        const float u = tvec.x * (-e2.z) + tvec.z * e2.x;
        if (u < 0.0f || u > det) continue;
        const V3 qvec = cross(tvec, e1);
        const float v = -qvec.y;
        if (v < 0.0f || u + v > det) continue;
        float dist = dot(e2, qvec);
        if (dist < 0.0f) continue;
        dist *= (1.0f / det);
        if (!(dist < length)) continue;
        if (best < 0.0f || dist < best) { best = dist; bestN = cross(e1, e2); bestS = surf; }
    }
    if (best >= 0.0f) {
        h.hit = 1; h.pos = v3(o.x + 0.0f * best, o.y + -1.0f * best, o.z + 0.0f * best);
        h.normal = norm(bestN); h.surface = bestS;
    }
    return h;
}

/* Track.cpp:469-494 getLineIntersection */
PD_HD bool line_intersection(float p0x, float p0y, float p1x, float p1y, float p2x, float p2y, float p3x, float p3y, float& ix, float& iy) {
    float s1x = p1x - p0x, s1y = p1y - p0y, s2x = p3x - p2x, s2y = p3y - p2y;
    float s = (-s1y * (p0x - p2x) + s1x * (p0y - p2y)) / (-s2x * s1y + s1x * s2y);
    float t = (s2x * (p0y - p2y) - s2y * (p0x - p2x)) / (-s2x * s1y + s1x * s2y);
    if (s >= 0 && s <= 1 && t >= 0 && t <= 1) { ix = p0x + (t * s1x); iy = p0y + (t * s1y); return true; }
    return false;
}

/* Cheap exact pre-filter for line_intersection: true when the full test is KNOWN to fail, decided from the numerators
 * and the denominator alone (the same three expressions line_intersection divides).  den == 0 makes s, t infinite or
 * NaN; a numerator of the opposite sign makes the quotient negative; |num| > |den| * (1 + 1e-6) makes it round to
 * more than 1.  Anything else (including NaNs: every comparison here is then false) goes through the full test. */
PD_HD bool line_intersection_rejects(float p0x, float p0y, float p1x, float p1y, float p2x, float p2y, float p3x, float p3y) {
    const float s1x = p1x - p0x, s1y = p1y - p0y, s2x = p3x - p2x, s2y = p3y - p2y;
    const float den = -s2x * s1y + s1x * s2y;
    const float ns = -s1y * (p0x - p2x) + s1x * (p0y - p2y);
    const float nt = s2x * (p0y - p2y) - s2y * (p0x - p2x);
    if (den == 0.0f) return true;
    if ((ns < 0.0f && den > 0.0f) || (ns > 0.0f && den < 0.0f)) return true;
    if ((nt < 0.0f && den > 0.0f) || (nt > 0.0f && den < 0.0f)) return true;
    const float lim = fabsf(den) * 1.000001f;
    return fabsf(ns) > lim || fabsf(nt) > lim;
}

struct Seg8 { float ax, az, bx, bz, bx0, by0, bz0, pad; };
PD_HD Seg8 load_seg8(const float* rec, int k) {
    Seg8 r;
#if defined(__CUDA_ARCH__)
    const float4 a = __ldg(reinterpret_cast<const float4*>(rec) + 2 * (size_t)k), b = __ldg(reinterpret_cast<const float4*>(rec) + 2 * (size_t)k + 1);
    r.ax = a.x; r.az = a.y; r.bx = a.z; r.bz = a.w; r.bx0 = b.x; r.by0 = b.y; r.bz0 = b.z; r.pad = b.w;
#else
    const float* p = rec + (size_t)k * 8;
    r.ax = p[0]; r.az = p[1]; r.bx = p[2]; r.bz = p[3]; r.bx0 = p[4]; r.by0 = p[5]; r.bz0 = p[6]; r.pad = p[7];
#endif
    return r;
}

/* One probe of Track::rayCastTrackBounds (Track.cpp:497-562): closest intersection of the 2-D segment
 * A -> B with the left / right boundary polylines of the NEAR points (|best - cachePos|^2 < nearRSq).
 * The reference tests every near point; here a grid walk (2-D DDA from A) visits only the cells the probe
 * crosses and stops once the best hit lies before the exit of the current cell.  Every candidate goes
 * through the same line_intersection arithmetic, so the result equals the exhaustive minimum.
 * Returns FLT_MAX when nothing is hit; returns false if the walk could not be used (start outside the grid). */
PD_HDN bool probe_walk(const TrackDev& T, float ax, float az, float bx, float bz, V3 cachePos, float nearRSq, float& bestOut) {
    const PdBoundGrid& G = T.segGrid;
    const int nFat = T.info.nFatPoints;
    float fx = (ax - G.ox) * G.invCell, fz = (az - G.oz) * G.invCell;
    int ix = (int)floorf(fx), iz = (int)floorf(fz);
    if (ix < 0 || iz < 0 || ix >= G.nx || iz >= G.nz) return false;
    const float dx = bx - ax, dz = bz - az;
    const float len = sqrtf(dx * dx + dz * dz);
    const int sx = dx > 0 ? 1 : -1, sz = dz > 0 ? 1 : -1;
    /* ray parameter (0..1) at which the walk leaves the current cell along x / z */
    const float tdx = dx != 0.0f ? fabsf(G.cell / dx) : FLT_MAX, tdz = dz != 0.0f ? fabsf(G.cell / dz) : FLT_MAX;
    float tmx = dx != 0.0f ? ((G.ox + (ix + (sx > 0 ? 1 : 0)) * G.cell) - ax) / dx : FLT_MAX;
    float tmz = dz != 0.0f ? ((G.oz + (iz + (sz > 0 ? 1 : 0)) * G.cell) - az) / dz : FLT_MAX;
    float best = FLT_MAX;
    const float margin = 0.1f;   /* metres: cells are padded by 0.05 m, rounding is far below that */
    /* (Measured on B200, round 2, both without effect on the probe phase -- it is a chain of dependent L2 round trips, one per
       cell, not arithmetic: (1) fetching the headers of the next 8 cells of the walk in one batch; (2) one padded box per cell and
       boundary side with a slab test in front of the side's segments, which removes ~3/4 of the segment tests.  The plain loop stays.) */
    for (int guard = 0; guard < 4096; ++guard) {
        const int c = iz * G.nx + ix;
        const int s0 = T.segStart[c], s1 = T.segStart[c + 1];
        /* records of the cell are contiguous; the next record is fetched while the current one is tested */
        Seg8 cur; if (s0 < s1) cur = load_seg8(T.segRec, s0);
        for (int k = s0; k < s1; ++k) {
            Seg8 nxt = cur; if (k + 1 < s1) nxt = load_seg8(T.segRec, k + 1);
            if (sqlen(cachePos - v3(cur.bx0, cur.by0, cur.bz0)) < nearRSq) {
                float jx, jz;
                if (!line_intersection_rejects(ax, az, bx, bz, cur.ax, cur.az, cur.bx, cur.bz) &&
                    line_intersection(ax, az, bx, bz, cur.ax, cur.az, cur.bx, cur.bz, jx, jz)) { const float ex = ax - jx, ez = az - jz; best = tminf(best, sqrtf(ex * ex + ez * ez)); }
            }
            cur = nxt;
        }
        const float tExit = tminf(tmx, tmz);
        if (tExit >= 1.0f) break;                                   /* B lies in this cell */
        if (best != FLT_MAX && best + margin < tExit * len) break;  /* nothing beyond the exit can be closer */
        if (tmx < tmz) { ix += sx; tmx += tdx; } else { iz += sz; tmz += tdz; }
        if (ix < 0 || iz < 0 || ix >= G.nx || iz >= G.nz) break;    /* left the grid: no boundary out there */
    }
    bestOut = best;
    return true;
}

/* Track::getPointIdAtLocation over the near set (Track.cpp:579-596): search of the point grid in growing square
 * blocks of cells around `pos` (3x3, 5x5, 7x7).  A block of radius R proves its best candidate to be the global
 * nearest as soon as that candidate is closer than the nearest point any cell OUTSIDE the block could hold
 * (R cells + the position's offset inside its own cell); candidates are ordered by (distance, id), as the
 * reference's ascending-id scan with a strict `>` orders them.  Cells beyond the grid hold no points.  Returns false
 * (caller falls back to the exhaustive scan) only when nothing is found within 3 cells. */
#define PD_NEAREST_MAX_RING 3
PD_HD void nearest_cell_scan(const TrackDev& T, int c, V3 pos, V3 cachePos, float nearRSq, float& bestDistSq, int& best) {
    const int k0 = T.ptStart[c], k1 = T.ptStart[c + 1];
    for (int k = k0; k < k1; ++k) {
#if defined(__CUDA_ARCH__)
        const float4 r = __ldg(reinterpret_cast<const float4*>(T.ptRec) + k);
        const V3 loc = v3(r.x, r.y, r.z); const int id = __float_as_int(r.w);
#else
        const float* r = T.ptRec + (size_t)k * 4;
        const V3 loc = v3(r[0], r[1], r[2]); int id; memcpy(&id, &r[3], 4);
#endif
        if (!(sqlen(cachePos - loc) < nearRSq)) continue;
        const float dsq = sqlen(loc - pos);
        if (bestDistSq > dsq || (bestDistSq == dsq && id < best)) { bestDistSq = dsq; best = id; }
    }
}
PD_HDN bool nearest_point_grid(const TrackDev& T, V3 pos, V3 cachePos, float nearRSq, int& bestPoint) {
    const PdBoundGrid& G = T.grid;
    const int cx = (int)floorf((pos.x - G.ox) * G.invCell), cz = (int)floorf((pos.z - G.oz) * G.invCell);
    if (cx < 0 || cz < 0 || cx >= G.nx || cz >= G.nz) return false;
    const float fx = (pos.x - G.ox) - cx * G.cell, fz = (pos.z - G.oz) - cz * G.cell;
    const float inCell = tminf(tminf(fx, G.cell - fx), tminf(fz, G.cell - fz));
    float bestDistSq = FLT_MAX; int best = -1;
    for (int R = 1; R <= PD_NEAREST_MAX_RING; ++R) {
        for (int iz = cz - R; iz <= cz + R; ++iz) {
            if (iz < 0 || iz >= G.nz) continue;
            for (int ix = cx - R; ix <= cx + R; ++ix) {
                if (ix < 0 || ix >= G.nx) continue;
                if (R > 1 && iz > cz - R && iz < cz + R && ix > cx - R && ix < cx + R) continue;   /* inner block: already scanned */
                nearest_cell_scan(T, iz * G.nx + ix, pos, cachePos, nearRSq, bestDistSq, best);
            }
        }
        const float border = inCell + (float)R * G.cell - 0.01f;
        if (best >= 0 && bestDistSq < border * border) { bestPoint = best; return true; }
    }
    return false;
}

/* Track::getPointIdAtDistance (Track.cpp:564-577) */
PD_HD int point_id_at_distance(const TrackDev& T, float distanceNorm) {
    const int n = T.info.nFatPoints;
    if (!n) return 0;
    if (distanceNorm < 0.0f) distanceNorm += 1.0f; else if (distanceNorm > 1.0f) distanceNorm -= 1.0f;
    return (int)(tclampf(distanceNorm, 0.0f, 1.0f) * (float)(n - 1));
}
PD_HD V3 track_direction_at_distance(const TrackDev& T, float distanceNorm) {
    const int id = point_id_at_distance(T, distanceNorm);
    if (id < T.info.nFatPoints) { const PdFatPoint& f = T.fat[id]; return v3(f.forwardDir[0], f.forwardDir[1], f.forwardDir[2]); }
    return v3(0, 0, 0);
}

/* Spline3d::find_nearest_point (Core/Spline3d.cpp:34-75) restricted as Track.cpp:609-629 calls it */
PD_HD bool spline_nearest(const TrackDev& T, V3 pos, int seg1, int seg2, int& outId, float& outDist) {
    int best_id = 0; float best_dist = FLT_MAX, spline_dist = 0; bool found = false;
    const int np = T.info.nSplineNodes;
    if (seg1 >= np) seg1 = 0;
    if (seg2 >= np) seg2 = 0;
    if (!seg1 && !seg2) seg2 = np;
    int id = seg1;
    while (id != seg2) {
        const float* p = T.splineXYZ + (size_t)id * 3;
        const float d = sqlen(pos - v3(p[0], p[1], p[2]));
        if (best_dist >= d) { best_dist = d; best_id = id; spline_dist = T.splineDist[id]; found = true; }
        ++id; if (id >= np) id = 0;
    }
    outId = best_id; outDist = spline_dist;
    return found;
}

/* ---- the same two searches with the work split over the four lanes of a quad (pd_quad.h) ----
 * Both return exactly what the sequential forms return: the spline search keeps the LAST minimum in sequence order
 * (`best_dist >= d`), so lanes take consecutive chunks and ties between lanes go to the later chunk; the point
 * search orders candidates by (distance, id), which does not depend on the visiting order at all. */
template <class Ex> PD_HD bool spline_nearest_quad(const TrackDev& T, V3 pos, int seg1, int seg2, Ex& ex, int& outId, float& outDist) {
    const int np = T.info.nSplineNodes;
    if (seg1 >= np) seg1 = 0;
    if (seg2 >= np) seg2 = 0;
    int cnt = seg2 - seg1; if (cnt < 0) cnt += np;
    if (!seg1 && !seg2) cnt = np;
    const int parts = 4 * ex.nhalf, part = ex.lane + 4 * ex.half;      /* a helper quad doubles the number of chunks; chunks stay in sequence order */
    const int chunk = (cnt + parts - 1) / parts;
    int k0 = part * chunk; if (k0 > cnt) k0 = cnt;
    const int k1 = (k0 + chunk < cnt) ? k0 + chunk : cnt;
    float bd = FLT_MAX; int bk = -1;
    int id = seg1 + k0; if (id >= np) id -= np;
    PD_UNROLL4
    for (int k = k0; k < k1; ++k) {
        const float* p = T.splineXYZ + (size_t)id * 3;
        const float d = sqlen(pos - v3(p[0], p[1], p[2]));
        if (bd >= d) { bd = d; bk = k; }
        ++id; if (id >= np) id = 0;
    }
    PD_UNROLL
    for (int off = 1; off <= 2; off <<= 1) {
        const float od = ex.get(bd, ex.lane ^ off); const int ok = ex.get(bk, ex.lane ^ off);
        if (ok >= 0 && (bk < 0 || od < bd || (od == bd && ok > bk))) { bd = od; bk = ok; }
    }
    if (ex.nhalf == 2) {
        const float od = ex.peer(bd); const int ok = ex.peer(bk);
        if (ok >= 0 && (bk < 0 || od < bd || (od == bd && ok > bk))) { bd = od; bk = ok; }
    }
    if (bk < 0) { outId = 0; outDist = 0; return false; }
    int best = seg1 + bk; if (best >= np) best -= np;
    outId = best; outDist = T.splineDist[best];
    return true;
}

template <class Ex> PD_HD bool nearest_point_grid_quad(const TrackDev& T, V3 pos, V3 cachePos, float nearRSq, Ex& ex, int& bestPoint) {
    const PdBoundGrid& G = T.grid;
    const int cx = (int)floorf((pos.x - G.ox) * G.invCell), cz = (int)floorf((pos.z - G.oz) * G.invCell);
    if (cx < 0 || cz < 0 || cx >= G.nx || cz >= G.nz) return false;
    const float fx = (pos.x - G.ox) - cx * G.cell, fz = (pos.z - G.oz) - cz * G.cell;
    const float inCell = tminf(tminf(fx, G.cell - fx), tminf(fz, G.cell - fz));
    float bestDistSq = FLT_MAX; int best = -1;
    for (int R = 1; R <= PD_NEAREST_MAX_RING; ++R) {
        const int side = 2 * R + 1;
        for (int cell = ex.lane + 4 * ex.half; cell < side * side; cell += 4 * ex.nhalf) {          /* the block's cells, dealt round-robin to the lanes (a helper quad doubles them) */
            const int dz = cell / side - R, dx = cell % side - R;
            if (R > 1 && dz > -R && dz < R && dx > -R && dx < R) continue; /* inner block: already scanned */
            const int iz = cz + dz, ix = cx + dx;
            if (iz < 0 || iz >= G.nz || ix < 0 || ix >= G.nx) continue;
            nearest_cell_scan(T, iz * G.nx + ix, pos, cachePos, nearRSq, bestDistSq, best);
        }
        float qd = bestDistSq; int qi = best;                               /* quad-wide best so far (own candidates stay per lane) */
        PD_UNROLL
        for (int off = 1; off <= 2; off <<= 1) {
            const float od = ex.get(qd, ex.lane ^ off); const int oi = ex.get(qi, ex.lane ^ off);
            if (oi >= 0 && (qi < 0 || qd > od || (qd == od && oi < qi))) { qd = od; qi = oi; }
        }
        if (ex.nhalf == 2) {
            const float od = ex.peer(qd); const int oi = ex.peer(qi);
            if (oi >= 0 && (qi < 0 || qd > od || (qd == od && oi < qi))) { qd = od; qi = oi; }
        }
        const float border = inCell + (float)R * G.cell - 0.01f;
        if (qi >= 0 && qd < border * border) { bestPoint = qi; return true; }
    }
    return false;
}

} // namespace pd
