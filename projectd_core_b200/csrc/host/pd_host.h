/*
 * pd_host.h -- host-side data loaders of the batched Car::step path (C++17, no CUDA, no torch).
 *
 * They read the reference's own on-disk formats, following the reference's init code and its quirks,
 * and produce the POD blocks the kernels consume (include/pd_params.h):
 *   INI files      Core/INIReader.cpp:31-106     (no trimming, first key wins, value keeps an inline ';')
 *   .lut curves    Core/Curve.cpp:128-168, inline curves :170-203
 *   surfaces.bin   Sim/Track.cpp:97-149, Sim/Surface.h:26-45 (58-byte packed header + verts + u16 indices)
 *   spline.bin / spline.cache   Sim/Track.h:13-25, Sim/Track.cpp:274-311
 *   B-spline nodes Core/Spline3d.cpp:79-162
 */
#pragma once
#include <map>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../../include/pd_params.h"

namespace pdh {

struct Error : public std::runtime_error { using std::runtime_error::runtime_error; };

/* ---- INI ---- */
struct Ini {
    std::map<std::string, std::map<std::string, std::string>> sections;
    bool ready = false;
    std::string filename;
    Ini() {}
    explicit Ini(const std::string& path) { load(path); }
    bool load(const std::string& path);
    bool hasSection(const std::string& s) const { return sections.find(s) != sections.end(); }
    bool hasKey(const std::string& s, const std::string& k) const;
    const std::string& getString(const std::string& s, const std::string& k) const;
    int getInt(const std::string& s, const std::string& k) const;
    float getFloat(const std::string& s, const std::string& k) const;
    void getFloat3(const std::string& s, const std::string& k, float* out3) const;
    bool tryGetInt(const std::string& s, const std::string& k, int& out) const;
    bool tryGetFloat(const std::string& s, const std::string& k, float& out) const;
    bool tryGetString(const std::string& s, const std::string& k, std::string& out) const;
    PdCurve getCurve(const std::string& s, const std::string& k) const;
};
bool file_exists(const std::string& path);
std::vector<std::string> split(const std::string& s, const std::string& delim);
float stof_ref(const std::string& s);   /* std::stof semantics */
int stoi_ref(const std::string& s);
bool load_curve(const std::string& path, PdCurve& out);      /* Curve::load */
bool parse_inline_curve(const std::string& str, PdCurve& out); /* Curve::parseInline */
void curve_add(PdCurve& c, float ref, float val);
float curve_value(const PdCurve& c, float ref);

/* ---- car ---- */
struct SetupVar {      /* Car/SetupManager.h:27-66 (only what setTune needs) */
    std::string name;
    float* fvalue = nullptr; double* dvalue = nullptr;
    float mult = 1, minV = -3.402823466e+38f, maxV = 3.402823466e+38f, step = 0.01f;
    int spinnerType = 3;   /* RawFloat */
    bool tunable = false;
    std::vector<float> spinnerValues;
};
struct CarModel {
    PdCarParams P;
    std::vector<SetupVar> setupVars;
    std::string dataPath;
    void setTune(const std::string& name, float value);     /* SetupManager::setTune; throws for reference variables this build does not support */
    void setRawTune(const std::string& name, float value);  /* SetupManager::setRawTune */
    void setScoringVar(const std::string& name, float value);
    float getScoringVar(const std::string& name) const;
};
/* Car::init and every component init it calls, for the demo-car topology (STRUT front, AXLE rear, 2WD) */
void load_car(const std::string& basePath, const std::string& model, CarModel& out);
extern const char* const kScoringVarNames[PD_NUM_SCORING_VARS];

/* ---- track ---- */
struct BvhNodeH { float bmin[3]; int32_t left; float bmax[3]; int32_t count; };
#ifndef PD_TRI_STRIDE
#define PD_TRI_STRIDE 12
#endif
/* Track::computeFatPoints' configuration (Sim/Track.h:83-89, spline.ini [SPLINE], Track.cpp:180-205) */
struct TraceConfig {
    int32_t traceSides = 0;
    float rayOffsetY = 20.0f, rayLength = 100.0f, sideMax = 10.0f, diffHeightMax = 0.01f, diffGripMax = 0.1f, step = 0.01f;
    int32_t nBadSectors = 0; uint32_t badSectors[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
struct TrackModel {
    PdTrackInfo info;
    std::vector<float> slim;          /* spline.bin: 5 floats per point {best.xyz, sides[2]} (Sim/Track.h:13-17) */
    TraceConfig trace;
    bool needFat = false;             /* spline.cache missing / stale / ignored on request: `fat` must be computed (on the GPU) before finish_track_points */
    bool closedLoop = false; float hashCellSize = 50.0f;
    std::vector<PdSurface> surfaces;
    std::vector<float> tris;          /* PD_TRI_STRIDE (12) floats per triangle, leaf order: v0, e1, e2, surface id bits, 0, 0 */
    std::vector<int32_t> triSurf;
    std::vector<BvhNodeH> nodes;
    std::vector<PdFatPoint> fat;
    std::vector<float> fatDist;       /* Track::fatPointDistances */
    std::vector<float> splineXYZ, splineDist;
    PdBoundGrid grid;                            /* 8 m cells: spline points (nearest-point search) */
    PdBoundGrid segGrid;                         /* 2 m cells (coarser on very large tracks): boundary segments (probes) */
    PdBoundGrid colGrid;                         /* x-z grid of triangle lists for vertical rays */
    std::vector<int32_t> colStart, colItems;     /* CSR per cell: triangle indices (leaf order of `tris`) */
    std::vector<int32_t> segStart, segItems;   /* CSR per cell: boundary segments, item = id * 2 + side (0 left, 1 right) */
    std::vector<int32_t> ptStart, ptItems;     /* CSR per cell: fat point ids (by `best`) */
    std::vector<float> segRec;                 /* 8 floats per segItems entry: ax, az, bx, bz, best.xyz of the owning point, 0 */
    std::vector<float> ptRec;                  /* 4 floats per ptItems entry: best.xyz, id (as int bits) */
    std::vector<float> triRaw;                 /* 9 floats per triangle (leaf order): v0, v1, v2 as stored in surfaces.bin */
    PdBoundGrid collGrid;                      /* x-z grid over all triangles for collision detection */
    std::vector<int32_t> collStart, collItems; /* CSR, two lists per cell: [2c] TRACK triangles, [2c+1] WALL triangles */
    std::vector<float> collPlane;              /* per collItems entry, 16 B: unit normal of the triangle + its plane offset (normal . v0); zero for a degenerate triangle */
    std::vector<float> collRec;                /* per collItems entry, 32 B: box min xyz, triangle index bits, box max xyz, 0; lists sorted by descending ymax */
    std::vector<float> collCell;               /* per cell, 32 B: track y min / max, wall y min / max, then (int bits) first TRACK entry, first WALL entry, end */
    std::vector<float> collY;                  /* per cell: y range of its TRACK triangles, y range of its WALL triangles */
};
/* recomputeFat: ignore spline.cache and leave `fat` to be regenerated (TrackModel::needFat) */
void load_track(const std::string& basePath, const std::string& name, TrackModel& out, bool recomputeFat = false);
/* synthetic track generator for config 4 (large mesh): closed loop of `nPoints` spline points, tessellated */
void make_synthetic_track(int targetTris, float lengthMeters, TrackModel& out);
void build_column_grid(TrackModel& out);
void build_bvh(const std::vector<float>& verts9, const std::vector<int32_t>& surf, TrackModel& out);
void finish_track_points(TrackModel& out, bool closedLoop, float cellSize);

} // namespace pdh
