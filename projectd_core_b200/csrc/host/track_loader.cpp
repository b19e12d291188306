/*
 * track_loader.cpp -- track assets -> GPU-ready arrays.
 *
 * surfaces.bin (Sim/Track.cpp:97-149, header Sim/Surface.h:26-45), spline.bin / spline.cache
 * (Sim/Track.h:13-25, Track.cpp:274-311), cfg/sim.ini [ENVIRONMENT] TRACK_GRIP / [VERTEX_HASH]
 * (Track.cpp:38-42,217-224), spline.ini CLOSED_LOOP (Track.cpp:155-158), the derived spline data of
 * Track::initTrackPoints (Track.cpp:178-272) and BSpline3d::init_from_array (Core/Spline3d.cpp:79-162).
 *
 * All static meshes (track + wall blobs) go into ONE bounding-volume hierarchy: the reference walks a
 * dSimpleSpace of 510 geoms per ray (PhysicsEngineODE.cpp:175-176) and keeps the minimum depth, which is
 * the closest hit over the union of the meshes.
 */
#include "pd_host.h"
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>

namespace pdh {

#pragma pack(push, 1)
struct BlobSurfaceHdr {
    uint32_t magic, numVertices, numIndices, sectorID, collisionCategory;
    float gripMod, damping, sinHeight, sinLength, granularity, dirtAdditiveK, vibrationGain, vibrationLength, wavPitchSpeed;
    uint8_t isValidTrack, isPitlane;
};
#pragma pack(pop)
static_assert(sizeof(BlobSurfaceHdr) == 58, "BlobSurface header is 58 bytes packed");

struct V3h { float x, y, z; };
static inline V3h sub(V3h a, V3h b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3h add(V3h a, V3h b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3h mulf(V3h a, float f) { return {a.x * f, a.y * f, a.z * f}; }
static inline V3h divf(V3h a, float f) { return {a.x / f, a.y / f, a.z / f}; }
static inline float lenh(V3h a) { return sqrtf(a.x * a.x + a.y * a.y + a.z * a.z); }

/* ---------------- BVH: top-down, median split on the longest centroid axis, <= 4 triangles per leaf ---------------- */
namespace {
struct Builder {
    const std::vector<float>& v9; const std::vector<int32_t>& surf;
    std::vector<uint32_t> order; std::vector<float> cx, cy, cz;
    std::vector<float> tmin, tmax;       /* per-triangle bounds */
    std::vector<BvhNodeH> nodes;
    Builder(const std::vector<float>& v, const std::vector<int32_t>& s) : v9(v), surf(s) {}
    void bounds(uint32_t lo, uint32_t hi, float* mn, float* mx) {
        for (int k = 0; k < 3; ++k) { mn[k] = 3.4e38f; mx[k] = -3.4e38f; }
        for (uint32_t i = lo; i < hi; ++i) { uint32_t t = order[i]; for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], tmin[t * 3 + k]); mx[k] = std::max(mx[k], tmax[t * 3 + k]); } }
    }
    int maxDepth = 0;
    /* median split on the widest centroid axis, leaves of <= 4 triangles: the tree is balanced, depth <= ceil(log2(n / 4)) + 1
       (21 for a million triangles).  The traversal (pd_track.h ray_cast) pushes both children of an interior node after popping it,
       so its stack never holds more than depth + 1 entries: PD_RAY_STACK (48) cannot overflow; build_bvh checks it. */
    void build(int nodeIdx, uint32_t lo, uint32_t hi, int depth = 0) {
        if (depth > maxDepth) maxDepth = depth;
        float mn[3], mx[3]; bounds(lo, hi, mn, mx);
        for (int k = 0; k < 3; ++k) {   /* conservative padding: the box test is only a filter */
            float pad = 1e-4f * std::max(1.0f, std::max(fabsf(mn[k]), fabsf(mx[k])));
            nodes[nodeIdx].bmin[k] = mn[k] - pad; nodes[nodeIdx].bmax[k] = mx[k] + pad;
        }
        const uint32_t n = hi - lo;
        if (n <= 4) { nodes[nodeIdx].left = (int32_t)lo; nodes[nodeIdx].count = (int32_t)n; return; }
        float cmn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, cmx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        for (uint32_t i = lo; i < hi; ++i) { uint32_t t = order[i]; float c[3] = {cx[t], cy[t], cz[t]}; for (int k = 0; k < 3; ++k) { cmn[k] = std::min(cmn[k], c[k]); cmx[k] = std::max(cmx[k], c[k]); } }
        int axis = 0; float ext = cmx[0] - cmn[0];
        if (cmx[1] - cmn[1] > ext) { axis = 1; ext = cmx[1] - cmn[1]; }
        if (cmx[2] - cmn[2] > ext) { axis = 2; ext = cmx[2] - cmn[2]; }
        const std::vector<float>& c = axis == 0 ? cx : axis == 1 ? cy : cz;
        const uint32_t mid = lo + n / 2;
        std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](uint32_t a, uint32_t b) { return c[a] < c[b]; });
        const int left = (int)nodes.size();
        nodes.push_back(BvhNodeH{}); nodes.push_back(BvhNodeH{});
        nodes[nodeIdx].left = left; nodes[nodeIdx].count = 0;
        build(left, lo, mid, depth + 1); build(left + 1, mid, hi, depth + 1);
    }
};
}

void build_bvh(const std::vector<float>& verts9, const std::vector<int32_t>& surf, TrackModel& out) {
    const uint32_t nt = (uint32_t)surf.size();
    Builder B(verts9, surf);
    B.order.resize(nt); std::iota(B.order.begin(), B.order.end(), 0u);
    B.cx.resize(nt); B.cy.resize(nt); B.cz.resize(nt); B.tmin.resize(nt * 3); B.tmax.resize(nt * 3);
    for (uint32_t t = 0; t < nt; ++t) {
        const float* p = &verts9[(size_t)t * 9];
        for (int k = 0; k < 3; ++k) {
            float a = p[k], b = p[3 + k], c = p[6 + k];
            B.tmin[t * 3 + k] = std::min(a, std::min(b, c)); B.tmax[t * 3 + k] = std::max(a, std::max(b, c));
        }
        B.cx[t] = (p[0] + p[3] + p[6]) / 3.0f; B.cy[t] = (p[1] + p[4] + p[7]) / 3.0f; B.cz[t] = (p[2] + p[5] + p[8]) / 3.0f;
    }
    B.nodes.reserve(nt); B.nodes.push_back(BvhNodeH{});
    if (nt > 0) B.build(0, 0, nt);
    if (B.maxDepth + 1 > 46) throw Error("BVH deeper than the ray caster's stack allows");      /* cannot happen with the median split; guards a future builder */
    out.nodes = B.nodes;
    out.tris.assign((size_t)nt * PD_TRI_STRIDE, 0.0f); out.triSurf.resize(nt);
    for (uint32_t i = 0; i < nt; ++i) {
        const uint32_t t = B.order[i]; const float* p = &verts9[(size_t)t * 9]; float* q = &out.tris[(size_t)i * PD_TRI_STRIDE];
        q[0] = p[0]; q[1] = p[1]; q[2] = p[2];
        q[3] = p[3] - p[0]; q[4] = p[4] - p[1]; q[5] = p[5] - p[2];     /* e1 = v1 - v0 */
        q[6] = p[6] - p[0]; q[7] = p[7] - p[1]; q[8] = p[8] - p[2];     /* e2 = v2 - v0 */
        out.triSurf[i] = surf[t];
        memcpy(&q[9], &surf[t], 4);                                      /* surface id inline (int bits) */
    }
    out.info.nTris = (int32_t)nt; out.info.nNodes = (int32_t)out.nodes.size();
    build_column_grid(out);
    /* collision detection (pd_collide.h): the triangles' own vertices (the separating-axis and edge tests start from the
       vertices, not from the rounded edge vectors of `tris`) and a 2 m x-z grid over ALL triangles, two lists per cell (TRACK, WALL) */
    out.triRaw.assign((size_t)nt * 9, 0.0f);
    for (uint32_t i = 0; i < nt; ++i) memcpy(&out.triRaw[(size_t)i * 9], &verts9[(size_t)B.order[i] * 9], 36);
    {
        PdBoundGrid& G = out.collGrid; memset(&G, 0, sizeof(G));
        out.collStart.assign(1, 0); out.collItems.clear();
        if (nt > 0) {
            float x0 = 3.4e38f, x1 = -3.4e38f, z0 = 3.4e38f, z1 = -3.4e38f;
            for (uint32_t i = 0; i < nt; ++i) for (int v = 0; v < 3; ++v) { const float* p = &out.triRaw[(size_t)i * 9 + v * 3]; x0 = std::min(x0, p[0]); x1 = std::max(x1, p[0]); z0 = std::min(z0, p[2]); z1 = std::max(z1, p[2]); }
            for (float cell = 2.0f;; cell *= 2.0f) {
                G.cell = cell; G.invCell = 1.0f / cell; G.ox = x0 - cell; G.oz = z0 - cell;
                G.nx = (int)ceilf((x1 - G.ox) * G.invCell) + 2; G.nz = (int)ceilf((z1 - G.oz) * G.invCell) + 2;
                if ((double)G.nx * G.nz <= 4.0e6 || cell >= 64.0f) break;
            }
            const size_t nc = (size_t)G.nx * G.nz;
            auto range = [&](uint32_t i, int& a, int& b, int& c, int& d) {
                const float* p = &out.triRaw[(size_t)i * 9];
                const float bx0 = std::min(p[0], std::min(p[3], p[6])) - 1e-3f, bx1 = std::max(p[0], std::max(p[3], p[6])) + 1e-3f;
                const float bz0 = std::min(p[2], std::min(p[5], p[8])) - 1e-3f, bz1 = std::max(p[2], std::max(p[5], p[8])) + 1e-3f;
                a = std::max(0, (int)floorf((bx0 - G.ox) * G.invCell)); b = std::min(G.nx - 1, (int)floorf((bx1 - G.ox) * G.invCell));
                c = std::max(0, (int)floorf((bz0 - G.oz) * G.invCell)); d = std::min(G.nz - 1, (int)floorf((bz1 - G.oz) * G.invCell));
            };
            auto category = [&](uint32_t i) { const int32_t sid = out.triSurf[i]; return (sid >= 0 && sid < (int32_t)out.surfaces.size()) ? (out.surfaces[sid].collisionCategory & 3u) : 0u; };
            /* per cell: TRACK triangles first, then WALL triangles (collStart[2c] .. collStart[2c+1] .. collStart[2c+2]), and the
               height range of each part (collY[4c..]: track min, track max, wall min, wall max) so that a whole cell is skipped
               when the collider's box does not reach its triangles */
            std::vector<int32_t> count(2 * nc + 1, 0);
            out.collY.assign(4 * nc, 0.0f);
            for (size_t c = 0; c < nc; ++c) { out.collY[4 * c] = 3.4e38f; out.collY[4 * c + 1] = -3.4e38f; out.collY[4 * c + 2] = 3.4e38f; out.collY[4 * c + 3] = -3.4e38f; }
            for (uint32_t i = 0; i < nt; ++i) {
                const uint32_t cat = category(i); if (cat != 1u && cat != 2u) continue;
                const float* p = &out.triRaw[(size_t)i * 9];
                const float y0 = std::min(p[1], std::min(p[4], p[7])), y1 = std::max(p[1], std::max(p[4], p[7]));
                int a2, b2, c2, d2; range(i, a2, b2, c2, d2);
                /* TRACK triangles that face up: the top of the cell's height range is the triangle's highest point INSIDE the cell's column
                   (the triangle clipped to the cell square), not its highest vertex -- a large road triangle on a slope reaches the floor
                   box's height only in cells the box does not stand on, and its entry in those cells is tested there.  The floor test's
                   triangle-normal axis is one-sided (box_tri_contact), so a triangle facing down keeps its full range. */
                const float nY = (p[5] - p[2]) * (p[6] - p[0]) - (p[3] - p[0]) * (p[8] - p[2]);       /* y of (v1 - v0) x (v2 - v0) */
                const bool clipTop = cat == 1u && nY > 0.0f && !getenv("PD_COLL_NO_CLIP");
                for (int iz = c2; iz <= d2; ++iz) for (int ix = a2; ix <= b2; ++ix) {
                    const size_t c = (size_t)iz * G.nx + ix;
                    count[2 * c + (cat - 1) + 1]++;
                    float top = y1;
                    if (clipTop && (a2 != b2 || c2 != d2)) {
                        /* Sutherland-Hodgman against the four sides of the cell square (x, z), grown by 1 mm; y interpolated along the edges */
                        float px[8][3], qx[8][3]; int np = 3;
                        for (int v = 0; v < 3; ++v) { px[v][0] = p[3 * v]; px[v][1] = p[3 * v + 1]; px[v][2] = p[3 * v + 2]; }
                        const float lim[4] = {G.ox + ix * G.cell - 1e-3f, G.ox + (ix + 1) * G.cell + 1e-3f, G.oz + iz * G.cell - 1e-3f, G.oz + (iz + 1) * G.cell + 1e-3f};
                        for (int side = 0; side < 4 && np > 0; ++side) {
                            const int ax = side < 2 ? 0 : 2; const float L = lim[side]; const bool keepGreater = (side & 1) == 0;
                            int nq = 0;
                            for (int v = 0; v < np; ++v) {
                                const float* A = px[v]; const float* B = px[(v + 1) % np];
                                const bool inA = keepGreater ? A[ax] >= L : A[ax] <= L, inB = keepGreater ? B[ax] >= L : B[ax] <= L;
                                if (inA) { memcpy(qx[nq++], A, 12); }
                                if (inA != inB) { const float t = (L - A[ax]) / (B[ax] - A[ax]); for (int d = 0; d < 3; ++d) qx[nq][d] = A[d] + t * (B[d] - A[d]); ++nq; }
                            }
                            np = nq; memcpy(px, qx, sizeof(float) * 3 * (size_t)np);
                        }
                        if (np > 0) { top = -3.4e38f; for (int v = 0; v < np; ++v) top = std::max(top, px[v][1]); top = std::min(y1, top + 1e-3f + 1e-5f * fabsf(top)); }
                        else top = y0;        /* the triangle's box reaches the cell, the triangle does not: it cannot be met here */
                    }
                    float* yr = &out.collY[4 * c + 2 * (cat - 1)]; yr[0] = std::min(yr[0], y0); yr[1] = std::max(yr[1], top);
                }
            }
            for (size_t c = 0; c < 2 * nc; ++c) count[c + 1] += count[c];
            out.collStart = count; out.collItems.assign((size_t)count[2 * nc], 0);
            std::vector<int32_t> fill(count.begin(), count.end() - 1);
            for (uint32_t i = 0; i < nt; ++i) {
                const uint32_t cat = category(i); if (cat != 1u && cat != 2u) continue;
                int a2, b2, c2, d2; range(i, a2, b2, c2, d2);
                for (int iz = c2; iz <= d2; ++iz) for (int ix = a2; ix <= b2; ++ix) out.collItems[(size_t)fill[2 * ((size_t)iz * G.nx + ix) + (cat - 1)]++] = (int32_t)i;
            }
            /* entries carry the triangle's own box, so that most are dismissed without touching the triangle; every list is
               sorted by descending top (ymax): a collider whose underside is above an entry's top is above all that follow */
            auto ymax_of = [&](int32_t i) { const float* p = &out.triRaw[(size_t)i * 9]; return std::max(p[1], std::max(p[4], p[7])); };
            for (size_t l = 0; l < 2 * nc; ++l) std::stable_sort(out.collItems.begin() + count[l], out.collItems.begin() + count[l + 1], [&](int32_t x, int32_t y) { return ymax_of(x) > ymax_of(y); });
            out.collRec.assign(out.collItems.size() * 8, 0.0f);
            out.collPlane.assign(out.collItems.size() * 4, 0.0f);
            for (size_t k = 0; k < out.collItems.size(); ++k) {
                const int32_t i = out.collItems[k]; const float* p = &out.triRaw[(size_t)i * 9]; float* q = &out.collRec[k * 8];
                for (int d = 0; d < 3; ++d) { q[d] = std::min(p[d], std::min(p[3 + d], p[6 + d])); q[4 + d] = std::max(p[d], std::max(p[3 + d], p[6 + d])); }
                memcpy(&q[3], &i, 4);
                /* the triangle's plane, for the one-sided plane test ahead of everything else (pd_collide.h): N = (v1 - v0) x (v2 - v0) as box_tri_contact forms it */
                const double e1[3] = {(double)p[3] - p[0], (double)p[4] - p[1], (double)p[5] - p[2]}, e2[3] = {(double)p[6] - p[0], (double)p[7] - p[1], (double)p[8] - p[2]};
                const double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                const double len = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                float* pl = &out.collPlane[k * 4];
                if (len > 1e-9) { pl[0] = (float)(n[0] / len); pl[1] = (float)(n[1] / len); pl[2] = (float)(n[2] / len); pl[3] = (float)((n[0] * p[0] + n[1] * p[1] + n[2] * p[2]) / len); }
            }
            /* one 32-byte header per cell: y ranges of its two lists and their bounds in collRec */
            out.collCell.assign(nc * 8, 0.0f);
            for (size_t c = 0; c < nc; ++c) {
                float* q = &out.collCell[c * 8];
                q[0] = out.collY[4 * c]; q[1] = out.collY[4 * c + 1]; q[2] = out.collY[4 * c + 2]; q[3] = out.collY[4 * c + 3];
                const int32_t k[3] = {count[2 * c], count[2 * c + 1], count[2 * c + 2]};
                memcpy(&q[4], k, 12);
            }
            if (getenv("PD_TRACK_STATS")) fprintf(stderr, "[pd] collision grid: cell %.1f m, %d x %d cells, %zu entries\n", G.cell, G.nx, G.nz, out.collItems.size());
        }
    }
}

/* Index for VERTICAL rays (every ray of the hot path is one: wheel rays Tyre.cpp:478-481, the teleport ray
 * Car.cpp:1243): uniform x-z grid, each cell lists the triangles whose padded x-z box touches it.
 * Triangles that no downward ray can hit are left out: the culling test of ray_cast_down rejects a triangle when
 * det = e1.x * (-e2.z) + e1.z * e2.x < 1e-6, and det does not depend on the ray -- evaluated here with the kernel's
 * own expression (no contraction), so the lists lose exactly the vertical (walls) and downward-facing triangles and
 * the hits are unchanged.  Cell size: the smallest of 0.5 / 1 / 2 / 4 ... m whose grid stays below 8M cells and
 * whose box-overlap count stays below ~24 entries per listed triangle; inside its box a triangle is listed only in
 * the cells its x-z projection really touches (separating-axis test against the padded cell, in double precision
 * with a 1 mm margin: conservative, a ray can only hit a triangle whose projection contains the ray's x-z point). */
static bool tri_touches_cell(const double tx[3], const double tz[3], double cx0, double cz0, double cx1, double cz1) {
    for (int i = 0; i < 3; ++i) {
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        const double nx = -(tz[j] - tz[i]), nz = tx[j] - tx[i];            /* edge normal */
        const double nl = std::sqrt(nx * nx + nz * nz);
        if (nl == 0.0) continue;
        const double d0 = nx * tx[i] + nz * tz[i], dk = nx * tx[k] + nz * tz[k];
        const double tmin = std::min(d0, dk), tmax = std::max(d0, dk);     /* the triangle's extent along the normal */
        const double b0 = nx * cx0 + nz * cz0, b1 = nx * cx1 + nz * cz0, b2 = nx * cx0 + nz * cz1, b3 = nx * cx1 + nz * cz1;
        const double bmin = std::min(std::min(b0, b1), std::min(b2, b3)), bmax = std::max(std::max(b0, b1), std::max(b2, b3));
        const double slack = 1e-3 * nl;
        if (bmin > tmax + slack || bmax < tmin - slack) return false;
    }
    return true;
}
void build_column_grid(TrackModel& out) {
    const size_t nt = out.triSurf.size();
    PdBoundGrid& G = out.colGrid; memset(&G, 0, sizeof(G));
    out.colStart.assign(1, 0); out.colItems.clear();
    if (nt == 0) return;
    float x0 = 3.4e38f, x1 = -3.4e38f, z0 = 3.4e38f, z1 = -3.4e38f;
    std::vector<float> bx0(nt), bx1(nt), bz0(nt), bz1(nt);
    std::vector<uint8_t> keep(nt, 0);
    size_t kept = 0;
    for (size_t t = 0; t < nt; ++t) {
        const float* p = &out.tris[t * PD_TRI_STRIDE];
        const float ax = p[0], az = p[2], bx = p[0] + p[3], bz = p[2] + p[5], cx = p[0] + p[6], cz = p[2] + p[8];
        bx0[t] = std::min(ax, std::min(bx, cx)) - 1e-3f; bx1[t] = std::max(ax, std::max(bx, cx)) + 1e-3f;
        bz0[t] = std::min(az, std::min(bz, cz)) - 1e-3f; bz1[t] = std::max(az, std::max(bz, cz)) + 1e-3f;
        x0 = std::min(x0, bx0[t]); x1 = std::max(x1, bx1[t]); z0 = std::min(z0, bz0[t]); z1 = std::max(z1, bz1[t]);
        const volatile float e1x = p[3], e1z = p[5], e2x = p[6], e2z = p[8];
        const volatile float m1 = e1x * (-e2z), m2 = e1z * e2x;
        const float det = m1 + m2;
        if (!(det < 0.000001f)) { keep[t] = 1; ++kept; }
    }
    for (float cell = 0.5f;; cell *= 2.0f) {
        G.cell = cell; G.invCell = 1.0f / cell; G.ox = x0 - cell; G.oz = z0 - cell;
        G.nx = (int)ceilf((x1 - G.ox) * G.invCell) + 2; G.nz = (int)ceilf((z1 - G.oz) * G.invCell) + 2;
        const double cells = (double)G.nx * G.nz;
        double pairs = 0;
        for (size_t t = 0; t < nt; ++t) if (keep[t]) pairs += (double)((int)floorf((bx1[t] - G.ox) * G.invCell) - (int)floorf((bx0[t] - G.ox) * G.invCell) + 1) * ((int)floorf((bz1[t] - G.oz) * G.invCell) - (int)floorf((bz0[t] - G.oz) * G.invCell) + 1);
        if ((pairs <= 24.0 * (double)std::max<size_t>(kept, 1) && cells <= 8.0e6) || cell >= 64.0f) break;
    }
    const size_t nc = (size_t)G.nx * G.nz;
    std::vector<int32_t> count(nc + 1, 0);
    auto range = [&](size_t t, int& ix0, int& ix1, int& iz0, int& iz1) {
        ix0 = std::max(0, (int)floorf((bx0[t] - G.ox) * G.invCell)); ix1 = std::min(G.nx - 1, (int)floorf((bx1[t] - G.ox) * G.invCell));
        iz0 = std::max(0, (int)floorf((bz0[t] - G.oz) * G.invCell)); iz1 = std::min(G.nz - 1, (int)floorf((bz1[t] - G.oz) * G.invCell));
    };
    auto touches = [&](size_t t, int ix, int iz) {
        const float* p = &out.tris[t * PD_TRI_STRIDE];
        const double tx[3] = {p[0], (double)p[0] + p[3], (double)p[0] + p[6]}, tz[3] = {p[2], (double)p[2] + p[5], (double)p[2] + p[8]};
        const double cx0 = (double)G.ox + (double)ix * G.cell - 1e-3, cz0 = (double)G.oz + (double)iz * G.cell - 1e-3;
        return tri_touches_cell(tx, tz, cx0, cz0, cx0 + G.cell + 2e-3, cz0 + G.cell + 2e-3);
    };
    for (size_t t = 0; t < nt; ++t) { if (!keep[t]) continue; int a, b, c, d; range(t, a, b, c, d); for (int iz = c; iz <= d; ++iz) for (int ix = a; ix <= b; ++ix) if (touches(t, ix, iz)) count[(size_t)iz * G.nx + ix + 1]++; }
    for (size_t c = 0; c < nc; ++c) count[c + 1] += count[c];
    out.colStart = count;
    out.colItems.assign((size_t)count[nc], 0);
    std::vector<int32_t> fill(count.begin(), count.end() - 1);
    for (size_t t = 0; t < nt; ++t) { if (!keep[t]) continue; int a, b, c, d; range(t, a, b, c, d); for (int iz = c; iz <= d; ++iz) for (int ix = a; ix <= b; ++ix) if (touches(t, ix, iz)) out.colItems[(size_t)fill[(size_t)iz * G.nx + ix]++] = (int32_t)t; }
    if (getenv("PD_TRACK_STATS")) {
        size_t occ = 0, mx = 0; for (size_t c = 0; c < nc; ++c) { const size_t k = (size_t)(count[c + 1] - count[c]); if (k) { ++occ; mx = std::max(mx, k); } }
        fprintf(stderr, "[pd] column grid: %zu of %zu triangles can face a downward ray; cell %.2f m, %d x %d cells (%zu occupied), %zu entries, mean %.1f / max %zu per occupied cell\n",
                kept, nt, G.cell, G.nx, G.nz, occ, out.colItems.size(), occ ? (double)out.colItems.size() / occ : 0.0, mx);
    }
}

/* BSpline3d::interpolate (Core/Spline3d.cpp:151-160), same operation order */
static V3h bspline(float u, V3h P0, V3h P1, V3h P2, V3h P3) {
    V3h point;
    point = divf(mulf(add(sub(add(mulf(P0, -1.0f), mulf(P1, 3.0f)), mulf(P2, 3.0f)), P3), u * u * u), 6.0f);
    point = add(point, divf(mulf(add(sub(mulf(P0, 3.0f), mulf(P1, 6.0f)), mulf(P2, 3.0f)), u * u), 6.0f));
    point = add(point, divf(mulf(add(mulf(P0, -3.0f), mulf(P2, 3.0f)), u), 6.0f));
    point = add(point, divf(add(add(P0, mulf(P1, 4.0f)), P2), 6.0f));
    return point;
}

/* index over boundary segments and spline points (see PdBoundGrid) */
static void build_bound_grid(TrackModel& out) {
    const int n = (int)out.fat.size();
    PdBoundGrid& G = out.grid; memset(&G, 0, sizeof(G));
    PdBoundGrid& S = out.segGrid; memset(&S, 0, sizeof(S));
    out.segStart.assign(1, 0); out.segItems.clear(); out.ptStart.assign(1, 0); out.ptItems.clear();
    if (n == 0) return;
    float x0 = 3.4e38f, x1 = -3.4e38f, z0 = 3.4e38f, z1 = -3.4e38f;
    for (const PdFatPoint& f : out.fat)
        for (const float* p : {f.best, f.left, f.right}) { x0 = std::min(x0, p[0]); x1 = std::max(x1, p[0]); z0 = std::min(z0, p[2]); z1 = std::max(z1, p[2]); }
    auto shape = [&](PdBoundGrid& g, float cell) {
        g.cell = cell; g.invCell = 1.0f / cell;
        const float m = std::max(2.0f * cell, 16.0f);          /* indexed margin around the track: cars further out fall back to the exhaustive scan */
        g.ox = x0 - m; g.oz = z0 - m;
        g.nx = (int)ceilf((x1 + m - g.ox) * g.invCell) + 1; g.nz = (int)ceilf((z1 + m - g.oz) * g.invCell) + 1;
    };
    /* point grid and segment grid: 8 m cells.  (Measured on B200: 2 m segment cells -- an empty corridor between the boundary
       polylines, 4x fewer segment tests -- made the probes SLOWER, 47 k instead of 36 k cycles per tick: a probe walk is a chain
       of dependent L2 loads, one per cell, and the finer grid quadruples the chain.)  Cells grow for very large tracks so that
       the index stays within a few million entries. */
    shape(G, 8.0f);
    { float cell = 8.0f; if (const char* q = getenv("PD_SEG_CELL")) { const float v = (float)atof(q); if (v >= 1.0f && v <= 64.0f) cell = v; }      /* tuning knob */
      shape(S, cell); while ((size_t)S.nx * S.nz > (size_t)4 << 20) { cell *= 1.5f; shape(S, cell); } }
    const size_t nc = (size_t)G.nx * G.nz, ncs = (size_t)S.nx * S.nz;
    std::vector<std::vector<int32_t>> seg(ncs), pts(nc);
    auto cx = [&](const PdBoundGrid& g, float x) { return std::max(0, std::min(g.nx - 1, (int)floorf((x - g.ox) * g.invCell))); };
    auto cz = [&](const PdBoundGrid& g, float z) { return std::max(0, std::min(g.nz - 1, (int)floorf((z - g.oz) * g.invCell))); };
    const float pad = 0.05f;   /* segments are listed in every cell their padded box touches */
    for (int id = 0; id < n; ++id) {
        const PdFatPoint& f = out.fat[id]; const PdFatPoint& g = out.fat[id + 1 < n ? id + 1 : 0];
        for (int side = 0; side < 2; ++side) {
            const float* a = side ? f.right : f.left; const float* b = side ? g.right : g.left;
            const int ix0 = cx(S, std::min(a[0], b[0]) - pad), ix1 = cx(S, std::max(a[0], b[0]) + pad), iz0 = cz(S, std::min(a[2], b[2]) - pad), iz1 = cz(S, std::max(a[2], b[2]) + pad);
            for (int iz = iz0; iz <= iz1; ++iz) for (int ix = ix0; ix <= ix1; ++ix) {
                /* long segments (closing segment of an open track, coarse splines): keep only the cells the padded segment really crosses */
                if ((ix1 - ix0) + (iz1 - iz0) > 2) {
                    const float bx0 = S.ox + ix * S.cell - pad, bx1 = bx0 + S.cell + 2 * pad, bz0 = S.oz + iz * S.cell - pad, bz1 = bz0 + S.cell + 2 * pad;
                    const float dx = b[0] - a[0], dz = b[2] - a[2];
                    float t0 = 0.0f, t1 = 1.0f; bool hit = true;
                    auto clip = [&](float p, float q) { if (p == 0.0f) { if (q < 0.0f) hit = false; return; } const float r = q / p; if (p < 0.0f) { if (r > t1) hit = false; else if (r > t0) t0 = r; } else { if (r < t0) hit = false; else if (r < t1) t1 = r; } };
                    clip(-dx, a[0] - bx0); if (hit) clip(dx, bx1 - a[0]); if (hit) clip(-dz, a[2] - bz0); if (hit) clip(dz, bz1 - a[2]);
                    if (!hit) continue;
                }
                seg[(size_t)iz * S.nx + ix].push_back(id * 2 + side);
            }
        }
        pts[(size_t)cz(G, f.best[2]) * G.nx + cx(G, f.best[0])].push_back(id);
    }
    out.segStart.resize(ncs + 1); out.ptStart.resize(nc + 1);
    for (size_t c = 0; c < ncs; ++c) { out.segStart[c] = (int32_t)out.segItems.size(); out.segItems.insert(out.segItems.end(), seg[c].begin(), seg[c].end()); }
    for (size_t c = 0; c < nc; ++c) { out.ptStart[c] = (int32_t)out.ptItems.size(); out.ptItems.insert(out.ptItems.end(), pts[c].begin(), pts[c].end()); }
    out.segStart[ncs] = (int32_t)out.segItems.size(); out.ptStart[nc] = (int32_t)out.ptItems.size();
    /* the same lists with the data inlined, so that a cell visit is index -> records (no id -> point indirection) */
    out.segRec.resize(out.segItems.size() * 8); out.ptRec.resize(out.ptItems.size() * 4);
    for (size_t k = 0; k < out.segItems.size(); ++k) {
        const int item = out.segItems[k], id = item >> 1;
        const PdFatPoint& f = out.fat[id]; const PdFatPoint& g = out.fat[id + 1 < n ? id + 1 : 0];
        const float* a = (item & 1) ? f.right : f.left; const float* b = (item & 1) ? g.right : g.left;
        float* r = &out.segRec[k * 8];
        r[0] = a[0]; r[1] = a[2]; r[2] = b[0]; r[3] = b[2]; r[4] = f.best[0]; r[5] = f.best[1]; r[6] = f.best[2]; r[7] = 0.0f;
    }
    for (size_t k = 0; k < out.ptItems.size(); ++k) {
        const int32_t id = out.ptItems[k]; const PdFatPoint& f = out.fat[id];
        float* r = &out.ptRec[k * 4];
        r[0] = f.best[0]; r[1] = f.best[1]; r[2] = f.best[2]; memcpy(&r[3], &id, 4);
    }
}

/* Track::initTrackPoints tail (Track.cpp:207-271) + BSpline3d::init_from_array */
void finish_track_points(TrackModel& out, bool closedLoop, float cellSize) {
    const int n = (int)out.fat.size();
    out.info.nFatPoints = n; out.info.closedLoop = closedLoop ? 1 : 0; out.info.hashCellSize = cellSize;
    float width = 0.1f, length = 0.1f;
    out.fatDist.assign(n, 0.0f);
    std::vector<V3h> pts(n);
    for (int id = 0; id < n; ++id) {
        const PdFatPoint& f = out.fat[id];
        pts[id] = {f.best[0], f.best[1], f.best[2]};
        const float w = lenh(sub({f.left[0], f.left[1], f.left[2]}, {f.right[0], f.right[1], f.right[2]}));
        if (width < w) width = w;
        out.fatDist[id] = length;
        if (id + 1 < n) { const PdFatPoint& g = out.fat[id + 1]; length += lenh(sub(pts[id], {g.best[0], g.best[1], g.best[2]})); }
    }
    out.info.computedTrackWidth = width; out.info.computedTrackLength = length;
    out.splineXYZ.clear(); out.splineDist.clear();
    if (n < 4) { out.info.nSplineNodes = 0; out.info.interpolateStep = 0; return; }
    const int steps = (int)(length / 0.1f) / n;
    out.info.interpolateStep = steps;
    std::vector<V3h> nodes;
    auto add_node = [&](V3h p) {
        nodes.push_back(p);
        if (nodes.size() == 1) out.splineDist.push_back(0.0f);
        else { const size_t k = nodes.size() - 1; out.splineDist.push_back(lenh(sub(nodes[k], nodes[k - 1])) + out.splineDist[k - 1]); }
    };
    auto wrap = [&](int id) { return id < n ? id : id - n; };
    if (!closedLoop) {
        const float d0 = lenh(sub(pts[1], pts[0])); const V3h n0 = divf(sub(pts[1], pts[0]), d0);
        for (int i = 0; i < steps; ++i) { float u = (float)i / (float)steps; add_node(add(pts[0], mulf(n0, u * d0))); }
    }
    for (int pt = 0; pt + 4 < n; ++pt)
        for (int i = 0; i < steps; ++i) { float u = (float)i / (float)steps; add_node(bspline(u, pts[pt], pts[pt + 1], pts[pt + 2], pts[pt + 3])); }
    if (closedLoop) {
        for (int pt = n - 4; pt < n; ++pt)
            for (int i = 0; i < steps; ++i) { float u = (float)i / (float)steps; add_node(bspline(u, pts[pt], pts[wrap(pt + 1)], pts[wrap(pt + 2)], pts[wrap(pt + 3)])); }
    } else {
        for (int pt = n - 3; pt + 1 < n; ++pt) {
            const float dx = lenh(sub(pts[pt + 1], pts[pt])); const V3h nx = divf(sub(pts[pt + 1], pts[pt]), dx);
            for (int i = 0; i < steps; ++i) { float u = (float)i / (float)steps; add_node(add(pts[pt], mulf(nx, u * dx))); }
        }
        add_node(pts[n - 1]);
    }
    out.splineXYZ.resize(nodes.size() * 3);
    for (size_t i = 0; i < nodes.size(); ++i) { out.splineXYZ[i * 3] = nodes[i].x; out.splineXYZ[i * 3 + 1] = nodes[i].y; out.splineXYZ[i * 3 + 2] = nodes[i].z; }
    out.info.nSplineNodes = (int32_t)nodes.size();
    out.info.computedTrackLength = out.splineDist.empty() ? length : out.splineDist.back();
    build_bound_grid(out);
}

static std::vector<uint8_t> read_file(const std::string& path) {
    std::vector<uint8_t> buf;
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return buf;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    if (n > 0 && fread(buf.data(), 1, (size_t)n, f) != (size_t)n) buf.clear();
    fclose(f);
    return buf;
}

void load_track(const std::string& basePathIn, const std::string& name, TrackModel& out, bool recomputeFat) {
    std::string basePath = basePathIn;
    if (!basePath.empty() && basePath.back() != '/') basePath += "/";
    const std::string folder = basePath + "content/tracks/" + name + "/";
    out = TrackModel(); memset(&out.info, 0, sizeof(out.info));
    out.info.dynamicGripLevel = 1.0f;
    float cellSize = 50.0f;
    {
        Ini sim(basePath + "cfg/sim.ini");
        if (sim.ready) { sim.tryGetFloat("ENVIRONMENT", "TRACK_GRIP", out.info.dynamicGripLevel); sim.tryGetFloat("VERTEX_HASH", "CELL_SIZE", cellSize); }
    }
    /* surfaces.bin */
    const std::vector<uint8_t> blob = read_file(folder + "surfaces.bin");
    if (blob.empty()) throw Error("cannot read " + folder + "surfaces.bin");
    std::vector<float> verts9; std::vector<int32_t> surf;
    size_t pos = 0;
    while (pos + sizeof(BlobSurfaceHdr) <= blob.size()) {
        BlobSurfaceHdr h; memcpy(&h, &blob[pos], sizeof(h)); pos += sizeof(h);
        if (h.magic != 0xAABBCCDD) throw Error("surfaces.bin: bad magic");
        if (!(h.numVertices > 0 && h.numIndices > 0)) throw Error("surfaces.bin: empty blob");
        const size_t vb = (size_t)h.numVertices * 12, ib = (size_t)h.numIndices * 2;
        if (pos + vb + ib > blob.size()) throw Error("surfaces.bin: truncated");
        const float* v = (const float*)&blob[pos]; const uint16_t* idx = (const uint16_t*)&blob[pos + vb];
        PdSurface s; memset(&s, 0, sizeof(s));
        s.gripMod = h.gripMod; s.damping = h.damping; s.sinHeight = h.sinHeight; s.sinLength = h.sinLength; s.granularity = h.granularity;
        s.dirtAdditiveK = h.dirtAdditiveK; s.collisionCategory = h.collisionCategory; s.sectorID = h.sectorID; s.isValidTrack = h.isValidTrack; s.isPitlane = h.isPitlane;
        const int32_t sid = (int32_t)out.surfaces.size();
        out.surfaces.push_back(s);
        std::vector<float> vcopy(v, v + (size_t)h.numVertices * 3);   /* unaligned-safe copy */
        for (uint32_t t = 0; t + 2 < h.numIndices; t += 3) {
            for (int k = 0; k < 3; ++k) { uint16_t vi; memcpy(&vi, (const uint8_t*)idx + (size_t)(t + k) * 2, 2); verts9.push_back(vcopy[(size_t)vi * 3]); verts9.push_back(vcopy[(size_t)vi * 3 + 1]); verts9.push_back(vcopy[(size_t)vi * 3 + 2]); }
            surf.push_back(sid);
        }
        pos += vb + ib;
    }
    out.info.nSurfaces = (int32_t)out.surfaces.size();
    build_bvh(verts9, surf, out);
    /* spline */
    bool closedLoop = false;
    {   /* Track::initTrackPoints (Track.cpp:178-205) */
        Ini sp(folder + "spline.ini");
        if (sp.ready) {
            closedLoop = sp.getInt("SPLINE", "CLOSED_LOOP") != 0;
            TraceConfig& tc = out.trace;
            int ts = 0; if (sp.tryGetInt("SPLINE", "TRACE_SIDES", ts)) tc.traceSides = ts != 0;
            sp.tryGetFloat("SPLINE", "TRACE_RAY_OFFSET_Y", tc.rayOffsetY); sp.tryGetFloat("SPLINE", "TRACE_RAY_LENGTH", tc.rayLength);
            sp.tryGetFloat("SPLINE", "TRACE_SIDE_MAX", tc.sideMax); sp.tryGetFloat("SPLINE", "TRACE_DIFF_HEIGHT_MAX", tc.diffHeightMax);
            sp.tryGetFloat("SPLINE", "TRACE_DIFF_GRIP_MAX", tc.diffGripMax); sp.tryGetFloat("SPLINE", "TRACE_STEP", tc.step);
            std::string list;
            if (sp.tryGetString("SPLINE", "TRACE_BAD_SECTORS", list)) for (const std::string& tok : split(list, "|")) if (!tok.empty() && tc.nBadSectors < 8) tc.badSectors[tc.nBadSectors++] = (uint32_t)stoi_ref(tok);
        }
    }
    out.closedLoop = closedLoop; out.hashCellSize = cellSize;
    const std::vector<uint8_t> slim = read_file(folder + "spline.bin");
    const std::vector<uint8_t> fat = read_file(folder + "spline.cache");
    const size_t nSlim = slim.size() / 20, nFat = fat.size() / sizeof(PdFatPoint);
    out.slim.resize(nSlim * 5); if (nSlim) memcpy(out.slim.data(), slim.data(), nSlim * 20);
    if (recomputeFat || nFat == 0 || nFat != nSlim) {
        /* Track::computeFatPoints (Track.cpp:366-433) is batch ray casting: it runs on the GPU (pd_batch.cu k_fat_points) once the
           triangle index is on the device; the point grids are built afterwards (finish_track_points) */
        if (nSlim == 0) throw Error("track '" + name + "': neither spline.cache nor spline.bin");
        out.needFat = true; out.fat.clear();
        return;
    }
    out.fat.resize(nFat); memcpy(out.fat.data(), fat.data(), nFat * sizeof(PdFatPoint));
    finish_track_points(out, closedLoop, cellSize);
}

/* Config 4: synthetic closed circuit, flat road strip + verges, tessellated to about `targetTris` triangles. */
void make_synthetic_track(int targetTris, float lengthMeters, TrackModel& out) {
    out = TrackModel(); memset(&out.info, 0, sizeof(out.info));
    out.info.dynamicGripLevel = 0.98f;
    const float R = lengthMeters / (2.0f * 3.14159265f);
    const float halfW = 6.0f, verge = 6.0f;
    const int across = 8;                                  /* quads across: 2 verge + 4 road + 2 verge */
    int along = std::max(64, targetTris / (across * 2));
    PdSurface road; memset(&road, 0, sizeof(road)); road.gripMod = 0.97f; road.collisionCategory = 1; road.isValidTrack = 1;
    PdSurface grass = road; grass.gripMod = 0.8f; grass.isValidTrack = 0; grass.dirtAdditiveK = 0.0f;
    out.surfaces.push_back(road); out.surfaces.push_back(grass);
    std::vector<float> verts9; std::vector<int32_t> surf;
    auto P = [&](int i, int j) {
        const float a = 2.0f * 3.14159265f * (float)(i % along) / (float)along;
        const float off = -(halfW + verge) + (2.0f * (halfW + verge)) * (float)j / (float)across;
        const float r = R * (1.0f + 0.25f * sinf(3.0f * a)) + off;
        return V3h{r * cosf(a), 2.0f * sinf(2.0f * a), r * sinf(a)};
    };
    for (int i = 0; i < along; ++i)
        for (int j = 0; j < across; ++j) {
            V3h a = P(i, j), b = P(i + 1, j), c = P(i + 1, j + 1), d = P(i, j + 1);
            const int32_t s = (j < 2 || j >= across - 2) ? 1 : 0;
            V3h t1[3] = {a, b, c}, t2[3] = {a, c, d};
            /* orient so that the geometric normal (v1-v0)x(v2-v0) points up (front face for a downward ray) */
            auto up = [&](V3h* t) { V3h e1 = sub(t[1], t[0]), e2 = sub(t[2], t[0]); float ny = e1.z * e2.x - e1.x * e2.z; if (ny < 0) std::swap(t[1], t[2]); };
            up(t1); up(t2);
            for (auto* t : {t1, t2}) { for (int k = 0; k < 3; ++k) { verts9.push_back(t[k].x); verts9.push_back(t[k].y); verts9.push_back(t[k].z); } surf.push_back(s); }
        }
    out.info.nSurfaces = 2;
    build_bvh(verts9, surf, out);
    const int nPts = std::max(16, (int)(lengthMeters / 1.5f));
    out.fat.resize(nPts);
    for (int i = 0; i < nPts; ++i) {
        auto C = [&](float a, float off) { const float r = R * (1.0f + 0.25f * sinf(3.0f * a)) + off; return V3h{r * cosf(a), 2.0f * sinf(2.0f * a), r * sinf(a)}; };
        const float a = 2.0f * 3.14159265f * (float)i / (float)nPts, a2 = 2.0f * 3.14159265f * (float)(i + 1) / (float)nPts;
        V3h c = C(a, 0), l = C(a, halfW), r = C(a, -halfW), nx = C(a2, 0);
        V3h f = sub(nx, c); const float fl = lenh(f); f = divf(f, fl);
        PdFatPoint& p = out.fat[i];
        p.best[0] = c.x; p.best[1] = c.y; p.best[2] = c.z; p.center[0] = c.x; p.center[1] = c.y; p.center[2] = c.z;
        p.left[0] = l.x; p.left[1] = l.y; p.left[2] = l.z; p.right[0] = r.x; p.right[1] = r.y; p.right[2] = r.z;
        p.forwardDir[0] = f.x; p.forwardDir[1] = f.y; p.forwardDir[2] = f.z;
    }
    finish_track_points(out, true, 50.0f);
}

} // namespace pdh
