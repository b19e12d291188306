/*
 * car_loader.cpp -- content/cars/<model>/data/ *.ini, *.lut  ->  PdCarParams (include/pd_params.h).
 *
 * Restates, for the demo-car topology (front STRUT, rear rigid AXLE, RWD/FWD, no turbo), the init code of
 *   Car::init / initCarData / initProbes / initLookAhead   Car/Car.cpp:31-316
 *   SuspensionStrut::init / attach / setPositions           Car/SuspensionStrut.cpp:20-223
 *   SuspensionAxle::init / attach                           Car/SuspensionAxle.cpp:16-113
 *   Tyre::initCompounds / setCompound                       Car/Tyre.cpp:48-392
 *   BrushSlipProvider / BrushTyreModel parameters           Car/Tyre.cpp:117-176
 *   Engine::init                                            Car/Engine.cpp:17-191
 *   Drivetrain::init                                        Car/Drivetrain.cpp:19-152
 *   AutoClutch::init, AutoBlip::init, AutoShifter::init     Car/AutoClutch.cpp:28-89, AutoBlip.cpp:17-49, AutoShifter.cpp:15-29
 *   BrakeSystem::init                                       Car/BrakeSystem.cpp:15-71
 *   AeroMap::init, Wing::init                               Car/AeroMap.cpp:16-81, Car/Wing.cpp:20-69
 *   SetupManager::init / setTune                            Car/SetupManager.cpp:18-215,283-399
 *   ScoringConfig::initDefaults                             Car/ScoringSystem.cpp:50-73
 *   Simulator::init                                         Sim/Simulator.cpp:22-90
 * and the joint set-up calls those make through Physics/ODE/JointODE.cpp:21-57 (anchors, slider axis,
 * relative rotations) with a minimal host-side body.  Anything the kernels do not implement is rejected
 * here, loudly, instead of being silently ignored.
 */
#include <cstdio>
#include <algorithm>
#include "pd_host.h"
#include <algorithm>
#include <cmath>
#include <cstring>

namespace pdh {

const char* const kScoringVarNames[PD_NUM_SCORING_VARS] = {
    "SmoothSteerSpeed", "MinBonusSpeed", "MaxBonusSpeed", "StallRpm", "DirectionThreshold", "OutOfTrackThreshold",
    "ApproachDistance", "CriticalDistance", "TravelBonus", "TravelSplineBonus", "DriftBonus", "SpeedBonus", "ThrottleBonus",
    "EngineRpmBonus", "DirectionBonus", "DirectionPenalty", "ObstApproachPenalty", "CollisionPenalty", "OffTrackPenalty",
    "GearGrindPenalty", "StallPenalty"};

/* ---- minimal host body (ODE dBody pose API) ---- */
namespace {
struct V { float x, y, z; };
inline V mk(const float* p) { return {p[0], p[1], p[2]}; }
inline V operator+(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V operator-(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V operator*(V a, float f) { return {a.x * f, a.y * f, a.z * f}; }
inline float dotv(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V crossv(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float lenv(V a) { return sqrtf(dotv(a, a)); }
inline V normv(V a) { float l = lenv(a); if (l != 0.0f) { float s = 1.0f / l; return a * s; } return a; }
inline void put(float* d, V v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

struct HBody {
    float pos[3] = {0, 0, 0};
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   /* ODE row-major */
    float q[4] = {1, 0, 0, 0};
    V toWorld(V p) const { return {R[0] * p.x + R[1] * p.y + R[2] * p.z + pos[0], R[3] * p.x + R[4] * p.y + R[5] * p.z + pos[1], R[6] * p.x + R[7] * p.y + R[8] * p.z + pos[2]}; }
    V toLocal(V p) const { V d = {p.x - pos[0], p.y - pos[1], p.z - pos[2]}; return {R[0] * d.x + R[3] * d.y + R[6] * d.z, R[1] * d.x + R[4] * d.y + R[7] * d.z, R[2] * d.x + R[5] * d.y + R[8] * d.z}; }
    V vecToLocal(V d) const { return {R[0] * d.x + R[3] * d.y + R[6] * d.z, R[1] * d.x + R[4] * d.y + R[7] * d.z, R[2] * d.x + R[5] * d.y + R[8] * d.z}; }
    void setPos(V p) { pos[0] = p.x; pos[1] = p.y; pos[2] = p.z; }
    /* dBodySetRotation from mat44f rows (axes) */
    void setRotationAxes(V ax, V ay, V az) {
        const float Rin[9] = {ax.x, ay.x, az.x, ax.y, ay.y, az.y, ax.z, ay.z, az.z};
        memcpy(R, Rin, sizeof(R));
        auto n3 = [](float* a) { float l = a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; if (l > 0) { l = 1.0f / sqrtf(l); a[0] *= l; a[1] *= l; a[2] *= l; } else { a[0] = 1; a[1] = 0; a[2] = 0; } };
        float n0 = R[0] * R[0] + R[1] * R[1] + R[2] * R[2]; if (n0 != 1.0f) n3(R);
        float proj = R[0] * R[3] + R[1] * R[4] + R[2] * R[5];
        if (proj != 0) { R[3] -= proj * R[0]; R[4] -= proj * R[1]; R[5] -= proj * R[2]; }
        float n1 = R[3] * R[3] + R[4] * R[4] + R[5] * R[5]; if (n1 != 1.0f) n3(R + 3);
        R[6] = R[1] * R[5] - R[2] * R[4]; R[7] = R[2] * R[3] - R[0] * R[5]; R[8] = R[0] * R[4] - R[1] * R[3];
        /* dRtoQ(raw input) + dNormalize4 */
        const float* M = Rin;
#define RR(i, j) M[(i) * 3 + (j)]
        float tr = RR(0, 0) + RR(1, 1) + RR(2, 2), s;
        if (tr >= 0) { s = sqrtf(tr + 1); q[0] = 0.5f * s; s = 0.5f * (1.0f / s); q[1] = (RR(2, 1) - RR(1, 2)) * s; q[2] = (RR(0, 2) - RR(2, 0)) * s; q[3] = (RR(1, 0) - RR(0, 1)) * s; }
        else {
            int c = 0; if (RR(1, 1) > RR(0, 0)) { c = (RR(2, 2) > RR(1, 1)) ? 2 : 1; } else if (RR(2, 2) > RR(0, 0)) c = 2;
            if (c == 0) { s = sqrtf((RR(0, 0) - (RR(1, 1) + RR(2, 2))) + 1); q[1] = 0.5f * s; s = 0.5f * (1.0f / s); q[2] = (RR(0, 1) + RR(1, 0)) * s; q[3] = (RR(2, 0) + RR(0, 2)) * s; q[0] = (RR(2, 1) - RR(1, 2)) * s; }
            else if (c == 1) { s = sqrtf((RR(1, 1) - (RR(2, 2) + RR(0, 0))) + 1); q[2] = 0.5f * s; s = 0.5f * (1.0f / s); q[3] = (RR(1, 2) + RR(2, 1)) * s; q[1] = (RR(0, 1) + RR(1, 0)) * s; q[0] = (RR(0, 2) - RR(2, 0)) * s; }
            else { s = sqrtf((RR(2, 2) - (RR(0, 0) + RR(1, 1))) + 1); q[3] = 0.5f * s; s = 0.5f * (1.0f / s); q[1] = (RR(2, 0) + RR(0, 2)) * s; q[2] = (RR(1, 2) + RR(2, 1)) * s; q[0] = (RR(1, 0) - RR(0, 1)) * s; }
        }
#undef RR
        float l = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
        if (l > 0) { l = 1.0f / sqrtf(l); for (int k = 0; k < 4; ++k) q[k] *= l; } else { q[0] = 1; q[1] = q[2] = q[3] = 0; }
    }
};
/* dQMultiply1: inv(b) * c */
inline void qmul1(float* a, const float* b, const float* c) {
    a[0] = b[0] * c[0] + b[1] * c[1] + b[2] * c[2] + b[3] * c[3];
    a[1] = b[0] * c[1] - b[1] * c[0] - b[2] * c[3] + b[3] * c[2];
    a[2] = b[0] * c[2] - b[2] * c[0] - b[3] * c[1] + b[1] * c[3];
    a[3] = b[0] * c[3] - b[3] * c[0] - b[1] * c[2] + b[2] * c[1];
}
inline void box_inertia(float m, float lx, float ly, float lz, float* I) { /* dMassSetBoxTotal */
    I[0] = m / 12.0f * (ly * ly + lz * lz); I[1] = m / 12.0f * (lx * lx + lz * lz); I[2] = m / 12.0f * (lx * lx + ly * ly);
}
/* dJointSetDBallAnchor1/2 with world points, distance = current */
inline void make_dball(PdDBall& d, const HBody& b0, const HBody& b1, V p1, V p2) {
    put(d.anchor1, b0.toLocal(p1)); put(d.anchor2, b1.toLocal(p2));
    V g1 = b0.toWorld(mk(d.anchor1)), g2 = b1.toWorld(mk(d.anchor2));
    d.distance = lenv(g1 - g2);
}
inline void load_damper(const Ini& ini, const std::string& id, PdDamper& d) {
    d.bumpSlow = ini.getFloat(id, "DAMP_BUMP"); d.reboundSlow = ini.getFloat(id, "DAMP_REBOUND");
    d.bumpFast = ini.getFloat(id, "DAMP_FAST_BUMP"); d.reboundFast = ini.getFloat(id, "DAMP_FAST_REBOUND");
    d.fastThresholdBump = ini.getFloat(id, "DAMP_FAST_BUMPTHRESHOLD"); d.fastThresholdRebound = ini.getFloat(id, "DAMP_FAST_REBOUNDTHRESHOLD");
    if (d.fastThresholdBump == 0.0f) d.fastThresholdBump = 0.2f;
    if (d.fastThresholdRebound == 0.0f) d.fastThresholdRebound = 0.2f;
    if (d.bumpFast == 0.0f) d.bumpFast = d.bumpSlow;
    if (d.reboundFast == 0.0f) d.reboundFast = d.reboundSlow;
}
inline float sgn(float x) { return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f); }
inline float load_sens_mult(float targetD, float targetLoad, float sensExp) { return (targetD * targetLoad) / powf(targetLoad, sensExp); }
}

static void load_strut(const Ini& ini, int index, const HBody& car, PdStrut& S, float& hubMassOut) {
    memset(&S, 0, sizeof(S));
    S.k = 90000.0f; S.baseCFM = 0.0000001f;
    const int iVer = ini.getInt("HEADER", "VERSION");
    const std::string id = index < 2 ? "FRONT" : "REAR";
    const float fWheelBase = ini.getFloat("BASIC", "WHEELBASE"), fCg = ini.getFloat("BASIC", "CG_LOCATION");
    const float fFrontBaseY = ini.getFloat("FRONT", "BASEY"), fFrontTrack = ini.getFloat("FRONT", "TRACK") * 0.5f;
    const float fRearBaseY = ini.getFloat("REAR", "BASEY"), fRearTrack = ini.getFloat("REAR", "TRACK") * 0.5f;
    V ref[4] = {{fFrontTrack, fFrontBaseY, (1.0f - fCg) * fWheelBase}, {-fFrontTrack, fFrontBaseY, (1.0f - fCg) * fWheelBase},
                {fRearTrack, fRearBaseY, -(fCg * fWheelBase)}, {-fRearTrack, fRearBaseY, -(fCg * fWheelBase)}};
    const V refPoint = ref[index];
    float t[3];
    auto g3 = [&](const char* k) { ini.getFloat3(id, k, t); return V{t[0], t[1], t[2]}; };
    V carStrut = g3("STRUT_CAR"), tyreStrut = g3("STRUT_TYRE"), wbF = g3("WBCAR_BOTTOM_FRONT"), wbR = g3("WBCAR_BOTTOM_REAR");
    V tyreWB = g3("WBTYRE_BOTTOM"), tyreSteer = g3("WBTYRE_STEER"), carSteer = g3("WBCAR_STEER");
    if (iVer >= 2) {
        const float rim = -ini.getFloat(id, "RIM_OFFSET");
        if (rim != 0.0f) { carStrut.x += rim; tyreStrut.x += rim; wbF.x += rim; wbR.x += rim; tyreWB.x += rim; tyreSteer.x += rim; carSteer.x += rim; }
    }
    const float hubMass = ini.getFloat(id, "HUB_MASS");
    S.bumpStopUp = ini.getFloat(id, "BUMPSTOP_UP"); S.bumpStopDn = -ini.getFloat(id, "BUMPSTOP_DN");
    S.rodLength = ini.getFloat(id, "ROD_LENGTH"); S.toeOutLinear = ini.getFloat(id, "TOE_OUT");
    S.k = ini.getFloat(id, "SPRING_RATE"); S.progressiveK = ini.getFloat(id, "PROGRESSIVE_SPRING_RATE");
    load_damper(ini, id, S.damper);
    S.bumpStopRate = ini.getFloat(id, "BUMP_STOP_RATE"); if (S.bumpStopRate == 0.0f) S.bumpStopRate = 500000.0f;
    S.staticCamber = -ini.getFloat(id, "STATIC_CAMBER") * 0.017453f; if (index % 2) S.staticCamber *= -1.0f;
    S.packerRange = ini.getFloat(id, "PACKER_RANGE");
    if (refPoint.x > 0.0f) { wbF.x *= -1.0f; wbR.x *= -1.0f; carSteer.x *= -1.0f; carStrut.x *= -1.0f; tyreWB.x *= -1.0f; tyreSteer.x *= -1.0f; tyreStrut.x *= -1.0f; }
    float fMass = hubMass; if (fMass <= 0.0f) fMass = 20.0f;
    hubMassOut = fMass * 0.8f;
    S.hubMass = fMass * 0.8f; box_inertia(S.hubMass, 0.2f, 0.6f, 0.6f, S.hubInertia);
    S.strutMass = fMass * 0.2f; box_inertia(S.strutMass, 0.05f, 0.5f, 0.2f, S.strutInertia);
    S.strutBodyLength = 0.2f;
    put(S.refPoint, refPoint); put(S.tyreStrut, tyreStrut); put(S.tyreSteer, tyreSteer);
    /* attach(): dataRelToBody.x = carBody->localToWorld(relToWheel.x + refPoint) with the chassis at the origin */
    const V bCarWB_F = car.toWorld(wbF + refPoint), bCarWB_R = car.toWorld(wbR + refPoint), bCarStrut = car.toWorld(carStrut + refPoint);
    const V bTyreWB = car.toWorld(tyreWB + refPoint), bCarSteer = car.toWorld(carSteer + refPoint), bTyreSteer = car.toWorld(tyreSteer + refPoint);
    put(S.carStrut, bCarStrut); put(S.baseCarSteer, bCarSteer);
    /* setPositions() */
    HBody hub, strut;
    hub.setRotationAxes({car.R[0], car.R[3], car.R[6]}, {car.R[1], car.R[4], car.R[7]}, {car.R[2], car.R[5], car.R[8]});
    hub.setPos(car.toWorld(refPoint));
    const V vCarStrut = car.toWorld(bCarStrut), vTyreStrut = hub.toWorld(tyreStrut);
    const V vNorm = normv(vTyreStrut - vCarStrut);
    const V vM3 = V{car.R[2], car.R[5], car.R[8]} * -1.0f;
    const V vM3N = crossv(vM3, vNorm);
    const V vM3NN = normv(crossv(vM3N, vNorm));
    strut.setRotationAxes(vM3NN, vM3N * -1.0f, vNorm * -1.0f);
    strut.setPos((vNorm * S.strutBodyLength) * 0.5f + vCarStrut);
    /* joints */
    make_dball(S.link[0], car, hub, bCarWB_R, bTyreWB);
    make_dball(S.link[1], car, hub, bCarWB_F, bTyreWB);
    make_dball(S.link[2], car, hub, bCarSteer, bTyreSteer);
    { /* init ends with setSteerLengthOffset(0): the steer rod is re-seated with the toe offset, target distance kept */
        const V cs = {bCarSteer.x + (0.0f + 0.0f + (sgn(refPoint.x) * S.toeOutLinear)), bCarSteer.y, bCarSteer.z};
        put(S.link[2].anchor1, car.toLocal(car.toWorld(cs))); put(S.link[2].anchor2, hub.toLocal(hub.toWorld(tyreSteer)));
    }
    { /* slider(strutBody, hub, axis = vTyreStrut - vCarStrut): dJointSetSliderAxis */
        const V axis = normv(vTyreStrut - vCarStrut);   /* setAxes normalises */
        put(S.sliderAxis1, strut.vecToLocal(axis));
        const V c = V{strut.pos[0] - hub.pos[0], strut.pos[1] - hub.pos[1], strut.pos[2] - hub.pos[2]};
        put(S.sliderOffset, hub.vecToLocal(c));
        qmul1(S.sliderQrel, strut.q, hub.q);
    }
    put(S.ballAnchor1, car.toLocal(vCarStrut)); put(S.ballAnchor2, strut.toLocal(vCarStrut));
    S.strutBaseLength = lenv(vTyreStrut - vCarStrut);
}

/* SuspensionDW::init + attach (SuspensionDW.cpp:18-195) for wheel `index` */
static void load_dw(const Ini& ini, int index, const HBody& car, PdDW& D) {
    memset(&D, 0, sizeof(D));
    D.k = 90000.0f; D.baseCFM = 0.0000001f;
    const int iVer = ini.getInt("HEADER", "VERSION");
    const std::string id = index < 2 ? "FRONT" : "REAR";
    const float fWheelBase = ini.getFloat("BASIC", "WHEELBASE"), fCg = ini.getFloat("BASIC", "CG_LOCATION");
    const float fFrontBaseY = ini.getFloat("FRONT", "BASEY"), fFrontTrack = ini.getFloat("FRONT", "TRACK") * 0.5f;
    const float fRearBaseY = ini.getFloat("REAR", "BASEY"), fRearTrack = ini.getFloat("REAR", "TRACK") * 0.5f;
    V ref[4] = {{fFrontTrack, fFrontBaseY, (1.0f - fCg) * fWheelBase}, {-fFrontTrack, fFrontBaseY, (1.0f - fCg) * fWheelBase},
                {fRearTrack, fRearBaseY, -(fCg * fWheelBase)}, {-fRearTrack, fRearBaseY, -(fCg * fWheelBase)}};
    const V refPoint = ref[index];
    float t[3];
    auto g3 = [&](const char* k) { ini.getFloat3(id, k, t); return V{t[0], t[1], t[2]}; };
    V carTopF = g3("WBCAR_TOP_FRONT"), carTopR = g3("WBCAR_TOP_REAR"), carBotF = g3("WBCAR_BOTTOM_FRONT"), carBotR = g3("WBCAR_BOTTOM_REAR");
    V tyreTop = g3("WBTYRE_TOP"), tyreBot = g3("WBTYRE_BOTTOM"), tyreSteer = g3("WBTYRE_STEER"), carSteer = g3("WBCAR_STEER");
    if (iVer >= 2) {
        const float rim = -ini.getFloat(id, "RIM_OFFSET");
        if (rim != 0.0f) { carTopF.x += rim; carTopR.x += rim; carBotF.x += rim; carBotR.x += rim; tyreTop.x += rim; tyreBot.x += rim; tyreSteer.x += rim; carSteer.x += rim; }
    }
    const float hubMass = ini.getFloat(id, "HUB_MASS");
    D.bumpStopUp = ini.getFloat(id, "BUMPSTOP_UP"); D.bumpStopDn = -ini.getFloat(id, "BUMPSTOP_DN");
    D.rodLength = ini.getFloat(id, "ROD_LENGTH"); D.toeOutLinear = ini.getFloat(id, "TOE_OUT");
    D.k = ini.getFloat(id, "SPRING_RATE"); D.progressiveK = ini.getFloat(id, "PROGRESSIVE_SPRING_RATE");
    load_damper(ini, id, D.damper);
    D.bumpStopRate = ini.getFloat(id, "BUMP_STOP_RATE"); if (D.bumpStopRate == 0.0f) D.bumpStopRate = 500000.0f;
    if (ini.hasKey(id, "BUMP_STOP_PROGRESSIVE")) D.bumpStopProgressive = ini.getFloat(id, "BUMP_STOP_PROGRESSIVE");
    D.staticCamber = -ini.getFloat(id, "STATIC_CAMBER") * 0.017453f; if (index % 2) D.staticCamber *= -1.0f;
    D.packerRange = ini.getFloat(id, "PACKER_RANGE");
    if (refPoint.x > 0.0f) { carBotF.x *= -1.0f; carBotR.x *= -1.0f; carSteer.x *= -1.0f; carTopF.x *= -1.0f; carTopR.x *= -1.0f; tyreBot.x *= -1.0f; tyreSteer.x *= -1.0f; tyreTop.x *= -1.0f; }
    float fMass = hubMass; if (fMass <= 0.0f) fMass = 20.0f;
    D.hubMass = fMass; box_inertia(fMass, 0.2f, 0.6f, 0.6f, D.hubInertia);      /* hubInertiaBox is never read from the ini: the default box */
    put(D.refPoint, refPoint); put(D.tyreSteer, tyreSteer);
    /* attach(): hub at the reference point with the chassis' rotation; dataRelToBody.x = carBody->localToWorld(relToWheel.x + refPoint) */
    HBody hub;
    hub.setRotationAxes({car.R[0], car.R[3], car.R[6]}, {car.R[1], car.R[4], car.R[7]}, {car.R[2], car.R[5], car.R[8]});
    hub.setPos(car.toWorld(refPoint));
    const V bCarBotF = car.toWorld(carBotF + refPoint), bCarBotR = car.toWorld(carBotR + refPoint), bCarTopF = car.toWorld(carTopF + refPoint), bCarTopR = car.toWorld(carTopR + refPoint);
    const V bTyreBot = car.toWorld(tyreBot + refPoint), bTyreTop = car.toWorld(tyreTop + refPoint), bCarSteer = car.toWorld(carSteer + refPoint), bTyreSteer = car.toWorld(tyreSteer + refPoint);
    make_dball(D.link[0], car, hub, bCarTopR, bTyreTop);
    make_dball(D.link[1], car, hub, bCarTopF, bTyreTop);
    make_dball(D.link[2], car, hub, bCarBotR, bTyreBot);
    make_dball(D.link[3], car, hub, bCarBotF, bTyreBot);
    make_dball(D.link[4], car, hub, bCarSteer, bTyreSteer);
    put(D.baseCarSteer, bCarSteer);
    { /* init ends with setSteerLengthOffset(0): the steer rod is re-seated with the toe offset, target distance kept */
        const V cs = {bCarSteer.x + (0.0f + 0.0f + (sgn(refPoint.x) * D.toeOutLinear)), bCarSteer.y, bCarSteer.z};
        put(D.link[4].anchor1, car.toLocal(car.toWorld(cs))); put(D.link[4].anchor2, hub.toLocal(hub.toWorld(tyreSteer)));
    }
}

/* SuspensionML::init (SuspensionML.cpp:15-99) for wheel `index`: kept in a PdDW (same bodies, joints and solver group), multilink = 1 */
static void load_ml(const Ini& ini, int index, const HBody& car, PdDW& D) {
    memset(&D, 0, sizeof(D));
    D.multilink = 1; D.baseCFM = 0.0000001f;      /* SuspensionML::SuspensionML; unused by its joints (setERPCFM is empty) */
    const std::string id = index < 2 ? "FRONT" : "REAR";
    const float fWheelBase = ini.getFloat("BASIC", "WHEELBASE"), fCg = ini.getFloat("BASIC", "CG_LOCATION");
    const float fFrontBaseY = ini.getFloat("FRONT", "BASEY"), fFrontTrack = ini.getFloat("FRONT", "TRACK") * 0.5f;
    const float fRearBaseY = ini.getFloat("REAR", "BASEY"), fRearTrack = ini.getFloat("REAR", "TRACK") * 0.5f;
    V ref[4] = {{fFrontTrack, fFrontBaseY, (1.0f - fCg) * fWheelBase}, {-fFrontTrack, fFrontBaseY, (1.0f - fCg) * fWheelBase},
                {fRearTrack, fRearBaseY, -(fCg * fWheelBase)}, {-fRearTrack, fRearBaseY, -(fCg * fWheelBase)}};
    const V basePosition = ref[index];
    put(D.refPoint, basePosition);
    D.hubMass = ini.getFloat(id, "HUB_MASS"); box_inertia(D.hubMass, 0.2f, 0.6f, 0.6f, D.hubInertia);
    HBody hub;
    hub.setRotationAxes({car.R[0], car.R[3], car.R[6]}, {car.R[1], car.R[4], car.R[7]}, {car.R[2], car.R[5], car.R[8]});
    hub.setPos(car.toWorld(basePosition));
    V tyre4 = {0, 0, 0};
    for (int i = 0; i < 5; ++i) {
        float t[3];
        ini.getFloat3(id, "JOINT" + std::to_string(i) + "_CAR", t); V ballCar = {t[0], t[1], t[2]};
        ini.getFloat3(id, "JOINT" + std::to_string(i) + "_TYRE", t); V ballTyre = {t[0], t[1], t[2]};
        if (basePosition.x > 0.0f) { ballCar.x *= -1.0f; ballTyre.x *= -1.0f; }
        const V carRel = car.toLocal(hub.toWorld(ballCar)), tyreRel = car.toLocal(hub.toWorld(ballTyre));
        make_dball(D.link[i], car, hub, car.toWorld(carRel), car.toWorld(tyreRel));
        if (i == 4) { put(D.baseCarSteer, carRel); tyre4 = ballTyre; }
    }
    put(D.tyreSteer, tyre4);
    D.rodLength = ini.getFloat(id, "ROD_LENGTH"); D.toeOutLinear = ini.getFloat(id, "TOE_OUT");
    D.k = ini.getFloat(id, "SPRING_RATE"); D.progressiveK = ini.getFloat(id, "PROGRESSIVE_SPRING_RATE");
    D.damper.bumpSlow = ini.getFloat(id, "DAMP_BUMP"); D.damper.reboundSlow = ini.getFloat(id, "DAMP_REBOUND");
    D.damper.bumpFast = ini.getFloat(id, "DAMP_FAST_BUMP"); D.damper.reboundFast = ini.getFloat(id, "DAMP_FAST_REBOUND");
    D.damper.fastThresholdBump = ini.getFloat(id, "DAMP_FAST_BUMPTHRESHOLD"); D.damper.fastThresholdRebound = ini.getFloat(id, "DAMP_FAST_REBOUNDTHRESHOLD");     /* no defaults here (SuspensionML.cpp:80-85) */
    D.staticCamber = -ini.getFloat(id, "STATIC_CAMBER") * 0.017453f; if (index % 2) D.staticCamber *= -1.0f;
}
/* HeaveSpring::init (HeaveSpring.cpp:11-54) */
static void load_heave(const Ini& ini, bool front, PdHeave& H) {
    memset(&H, 0, sizeof(H));
    const std::string id = front ? "HEAVE_FRONT" : "HEAVE_REAR";
    if (!ini.hasSection(id)) return;
    H.present = 1;
    H.bumpStopUp = ini.getFloat(id, "BUMPSTOP_UP"); H.bumpStopDn = -ini.getFloat(id, "BUMPSTOP_DN");
    H.rodLength = ini.getFloat(id, "ROD_LENGTH"); H.k = ini.getFloat(id, "SPRING_RATE"); H.progressiveK = ini.getFloat(id, "PROGRESSIVE_SPRING_RATE");
    load_damper(ini, id, H.damper);
    H.bumpStopRate = ini.getFloat(id, "BUMP_STOP_RATE"); if (H.bumpStopRate == 0.0f) H.bumpStopRate = 500000.0f;
    H.packerRange = ini.getFloat(id, "PACKER_RANGE");
}

static void load_axle(const Ini& ini, const HBody& car, PdAxle& A) {
    memset(&A, 0, sizeof(A));
    A.baseCFM = 0.0000001f; A.attachRelativePos = 1.0f;
    const int iVer = ini.getInt("HEADER", "VERSION");
    const float fWheelBase = ini.getFloat("BASIC", "WHEELBASE"), fCg = ini.getFloat("BASIC", "CG_LOCATION");
    A.referenceY = ini.getFloat("REAR", "BASEY"); A.track = ini.getFloat("REAR", "TRACK") * 0.5f;
    A.axleBasePos[0] = 0.0f; A.axleBasePos[1] = A.referenceY; A.axleBasePos[2] = -(fCg * fWheelBase);
    if (iVer >= 4) A.attachRelativePos = ini.getFloat("AXLE", "ATTACH_REL_POS");
    A.axleMass = ini.getFloat("REAR", "HUB_MASS"); box_inertia(A.axleMass, A.track * 2.0f, 0.2f, 0.5f, A.axleInertia);
    HBody axle;
    axle.setRotationAxes({car.R[0], car.R[3], car.R[6]}, {car.R[1], car.R[4], car.R[7]}, {car.R[2], car.R[5], car.R[8]});
    axle.setPos(car.toWorld(mk(A.axleBasePos)));
    A.nLinks = ini.getInt("AXLE", "LINK_COUNT");
    if (A.nLinks > PD_AXLE_LINKS || A.nLinks <= 0) throw Error("axle LINK_COUNT outside 1..5 is not supported");
    for (int i = 0; i < A.nLinks; ++i) {
        float c3[3], a3[3];
        ini.getFloat3("AXLE", "J" + std::to_string(i) + "_CAR", c3); ini.getFloat3("AXLE", "J" + std::to_string(i) + "_AXLE", a3);
        const V relCarBall = car.toLocal(axle.toWorld(mk(c3))), relAxleBall = car.toLocal(axle.toWorld(mk(a3)));
        make_dball(A.link[i], car, axle, car.toWorld(relCarBall), car.toWorld(relAxleBall));
    }
    A.bumpStopUp = ini.getFloat("REAR", "BUMPSTOP_UP"); A.bumpStopDn = -ini.getFloat("REAR", "BUMPSTOP_DN");
    A.rodLength = ini.getFloat("REAR", "ROD_LENGTH");
    A.k = ini.getFloat("REAR", "SPRING_RATE"); A.progressiveK = ini.getFloat("REAR", "PROGRESSIVE_SPRING_RATE");
    load_damper(ini, "REAR", A.damper);
    A.bumpStopRate = ini.getFloat("REAR", "BUMP_STOP_RATE"); if (A.bumpStopRate == 0.0f) A.bumpStopRate = 500000.0f;
    if (iVer >= 3) A.leafSpringKx = ini.getFloat("AXLE", "LEAF_SPRING_LAT_K");
    A.torqueReaction = ini.getFloat("AXLE", "TORQUE_REACTION");
}

static void load_tyre(const std::string& dataPath, int index, float ambient, PdTyre& T) {
    (void)ambient;
    memset(&T, 0, sizeof(T));
    Ini ini(dataPath + "tyres.ini");
    if (!ini.ready) throw Error("cannot read tyres.ini");
    const int iVer = ini.getInt("HEADER", "VERSION");
    if (iVer < 10) throw Error("tyres.ini VERSION < 10 is not supported (reference: GUARD_FATAL)");
    const std::string id = index < 2 ? "FRONT" : "REAR";
    /* defaults (Tyre.h:58-66, TyreCompound.h:9-76) */
    T.flatSpotK = 0.15f; T.explosionTemperature = 350.0f; T.pressureTemperatureGain = 0.16f; T.camberSpreadK = 1.4f;
    T.thermalFrictionK = 0.03f; T.thermalRollingK = 0.5f; T.blisterThreshold = 9000.0f; T.grainGamma = 1.0f; T.blisterGamma = 1.0f; T.optimumTemp = 80.0f;
    T.surfaceTransfer = 0.3f; T.patchTransfer = 0.2f; T.patchCoreTransfer = 0.2f; T.internalCoreTransfer = 0.004f;
    T.Dx1 = -0.0145f; T.brakeDXMod = 1.0f; T.cfXmult = 1.0f; T.dCamberBlend = 1.0f; T.pressureRRGain = 0.5f;
    if (ini.hasSection("EXPLOSION")) T.explosionTemperature = ini.getFloat("EXPLOSION", "TEMPERATURE");
    if (ini.hasSection("VIRTUALKM")) T.useLoadForVKM = ini.getInt("VIRTUALKM", "USE_LOAD") != 0;
    if (ini.hasSection("ADDITIONAL1")) {
        T.pressureTemperatureGain = ini.getFloat("ADDITIONAL1", "PRESSURE_TEMPERATURE_GAIN");
        const float sp = ini.getFloat("ADDITIONAL1", "CAMBER_TEMP_SPREAD_K"); if (sp != 0.0f) T.camberSpreadK = sp;
    }
    const std::string c = id;   /* compound 0 (cfg/demo.ini: TYRE_COMPOUND commented out; Tyre::init -> setCompound(0)) */
    if (!ini.hasSection(c)) throw Error("tyres.ini has no [" + c + "]");
    T.version = iVer;
    T.width = ini.getFloat(c, "WIDTH"); if (T.width <= 0) T.width = 0.15f;
    T.radius = ini.getFloat(c, "RADIUS"); T.rimRadius = ini.getFloat(c, "RIM_RADIUS");
    const float flexK = ini.getFloat(c, "FLEX"); (void)flexK;
    float fFLA = ini.getFloat(c, "FRICTION_LIMIT_ANGLE"); if (fFLA == 0.0f) fFLA = 7.5f;
    T.cfXmult = ini.getFloat(c, "CX_MULT");
    T.radiusRaiseK = ini.getFloat(c, "RADIUS_ANGULAR_K") * 0.001f;
    if (ini.hasKey(c, "BRAKE_DX_MOD")) { T.brakeDXMod = ini.getFloat(c, "BRAKE_DX_MOD"); if (T.brakeDXMod == 0.0f) T.brakeDXMod = 1.0f; else T.brakeDXMod += 1.0f; }
    if (ini.hasKey(c, "COMBINED_FACTOR")) T.combinedFactor = ini.getFloat(c, "COMBINED_FACTOR");
    const float fFZ0 = ini.getFloat(c, "FZ0"), fFlexGain = ini.getFloat(c, "FLEX_GAIN");
    T.lsExpX = ini.getFloat(c, "LS_EXPX"); T.lsExpY = ini.getFloat(c, "LS_EXPY");
    T.Dx0 = ini.getFloat(c, "DX_REF"); const float Dy0 = ini.getFloat(c, "DY_REF");
    T.lsMultX = load_sens_mult(T.Dx0, fFZ0, T.lsExpX); T.lsMultY = load_sens_mult(Dy0, fFZ0, T.lsExpY);
    if (ini.hasKey(c, "DY_CURVE") || ini.hasKey(c, "DX_CURVE") || ini.hasKey(c, "DCAMBER_LUT"))
        throw Error("tyre load / camber LUT curves (cubic-spline variants of SCTM) are not supported yet");
    T.sctmFz0 = fFZ0; T.Fz0 = 2000.0f;   /* TyreModelData::Fz0 keeps its default: initCompounds only sets the brush model's Fz0 */
    T.maxSlip0 = tanf(fFLA * 0.017453f); T.maxSlip1 = tanf(((fFlexGain + 1.0f) * fFLA) * 0.017453f);
    T.asy = ini.getFloat(c, "FALLOFF_LEVEL"); T.falloffSpeed = ini.getFloat(c, "FALLOFF_SPEED");
    T.speedSensitivity = ini.getFloat(c, "SPEED_SENSITIVITY"); T.relaxationLength = ini.getFloat(c, "RELAXATION_LENGTH");
    T.rr0 = ini.getFloat(c, "ROLLING_RESISTANCE_0"); T.rr1 = ini.getFloat(c, "ROLLING_RESISTANCE_1"); T.rr_slip = ini.getFloat(c, "ROLLING_RESISTANCE_SLIP");
    T.camberGain = ini.getFloat(c, "CAMBER_GAIN"); T.dcamber0 = ini.getFloat(c, "DCAMBER_0"); T.dcamber1 = ini.getFloat(c, "DCAMBER_1");
    if (T.dcamber0 == 0.0f || T.dcamber1 == 0.0f) { T.dcamber0 = 0.1f; T.dcamber1 = -0.8f; }
    T.angularInertia = ini.getFloat(c, "ANGULAR_INERTIA"); T.d = ini.getFloat(c, "DAMP"); T.k = ini.getFloat(c, "RATE");
    if (T.angularInertia == 0.0f) T.angularInertia = 1.2f;
    if (T.d == 0.0f) T.d = 400.0f;
    if (T.k == 0.0f) T.k = 220000.0f;
    if (T.Dx0 == 0.0f) T.Dx0 = Dy0 * 1.2f;
    if (T.Dx1 == 0.0f) T.Dx1 = -0.145f * 0.1f;
    T.pressureStaticDefault = ini.getFloat(c, "PRESSURE_STATIC"); if (T.pressureStaticDefault == 0.0f) T.pressureStaticDefault = 26.0f;
    T.pressureRef = T.pressureStaticDefault;
    T.pressureSpringGain = ini.getFloat(c, "PRESSURE_SPRING_GAIN"); if (T.pressureSpringGain == 0.0f) T.pressureSpringGain = 1000.0f;
    T.pressureCfGain = ini.getFloat(c, "PRESSURE_FLEX_GAIN");
    T.pressureRRGain = ini.getFloat(c, "PRESSURE_RR_GAIN"); T.pressureGainD = ini.getFloat(c, "PRESSURE_D_GAIN");
    T.idealPressure = ini.getFloat(c, "PRESSURE_IDEAL"); if (T.idealPressure == 0.0f) T.idealPressure = 26.0f;
    const std::string th = "THERMAL_" + c;
    if (ini.hasSection(th)) {
        T.surfaceTransfer = ini.getFloat(th, "SURFACE_TRANSFER"); T.patchTransfer = ini.getFloat(th, "PATCH_TRANSFER"); T.patchCoreTransfer = ini.getFloat(th, "CORE_TRANSFER");
        T.thermalFrictionK = ini.getFloat(th, "FRICTION_K"); T.thermalRollingK = ini.getFloat(th, "ROLLING_K");
        T.internalCoreTransfer = ini.getFloat(th, "INTERNAL_CORE_TRANSFER");
        if (ini.hasKey(th, "COOL_FACTOR")) T.coolFactorGain = (ini.getFloat(th, "COOL_FACTOR") - 1.0f) * 0.000324f;
        T.thermalRollingSurfaceK = ini.getFloat(th, "SURFACE_ROLLING_K");
        load_curve(dataPath + ini.getString(th, "PERFORMANCE_CURVE"), T.performanceCurve);
    }
    load_curve(dataPath + ini.getString(c, "WEAR_CURVE"), T.wearCurve);
    for (int i = 0; i < T.wearCurve.n; ++i) T.wearCurve.val[i] *= 0.01f;
    if (T.performanceCurve.n > 0) {
        for (int n = 0; n < T.performanceCurve.n; ++n) if (T.performanceCurve.val[n] >= 1.0f) { T.grainThreshold = T.performanceCurve.ref[n]; break; }
        for (int n = T.performanceCurve.n - 1; n > 0; --n) if (T.performanceCurve.val[n] >= 1.0f) { T.blisterThreshold = T.performanceCurve.ref[n]; T.optimumTemp = T.performanceCurve.ref[n]; break; }
    }
    T.blisterGamma = ini.getFloat(th, "BLISTER_GAMMA"); T.blisterGain = ini.getFloat(th, "BLISTER_GAIN");
    T.grainGamma = ini.getFloat(th, "GRAIN_GAMMA"); T.grainGain = ini.getFloat(th, "GRAIN_GAIN");
    {   /* loadSensExpD(lsExpY, lsMultY, 3000) */
        const float fSens = (powf(3000.0f, T.lsExpY) * T.lsMultY) / 3000.0f;
        T.softnessIndex = fSens - 1.0f > 0.0f ? fSens - 1.0f : 0.0f;
    }
    /* setCompound: SCTM copies */
    T.sctmLsMultX = T.lsMultX; T.sctmLsExpX = T.lsExpX;
}

static void load_engine(const std::string& dataPath, PdEngine& E, PdCarParams& P) {
    memset(&E, 0, sizeof(E));
    Ini ini(dataPath + "engine.ini");
    if (!ini.ready) throw Error("cannot read engine.ini");
    E.inertia = 1.0f; E.limiterMultiplier = 1.0f; E.bovThreshold = 0.2f; E.coast2 = 0.000001f; E.limiter = 18000; E.limiterCycles = 50;
    E.overlapFreq = 1.0f; E.overlapIdealRPM = 6000.0f;
    load_curve(dataPath + ini.getString("HEADER", "POWER_CURVE"), E.powerCurve);
    E.minimum = ini.getInt("ENGINE_DATA", "MINIMUM"); if (!E.minimum) E.minimum = 1000;
    if (ini.getString("HEADER", "COAST_CURVE") == "FROM_COAST_REF") {
        const float fRpm = ini.getFloat("COAST_REF", "RPM"), fTorque = ini.getFloat("COAST_REF", "TORQUE"), fNL = ini.getFloat("COAST_REF", "NON_LINEARITY");
        const float v13 = ((1.0f - fNL) * fRpm) - E.minimum, v14 = fNL * fRpm;
        E.coast1 = (v13 == 0.0f) ? 0.0f : -(fTorque / v13);
        E.coast2 = (v14 == 0.0f) ? 0.0f : fTorque / (v14 * v14);
    }
    E.inertia = ini.getFloat("ENGINE_DATA", "INERTIA");
    E.limiter = ini.getInt("ENGINE_DATA", "LIMITER");
    if (E.limiter) { E.rpmDamageThreshold = E.limiter * 1.05f; E.rpmDamageK = 10.0f; }
    E.limiterCycles = ini.getInt("ENGINE_DATA", "LIMITER_HZ");
    if (E.limiterCycles) E.limiterCycles = 1000 / E.limiterCycles / 3; else E.limiterCycles = 50;
    if (ini.hasSection("COAST_SETTINGS")) {
        PdCurve lut = ini.getCurve("COAST_SETTINGS", "LUT");
        const int def = ini.getInt("COAST_SETTINGS", "DEFAULT");
        if (def >= 0 && def < lut.n) E.gasCoastOffset = curve_value(lut, (float)def);
        E.coastEntryRpm = E.minimum + ini.getInt("COAST_SETTINGS", "ACTIVATION_RPM");
    }
    { /* Engine.cpp:69-94: TURBO_n sections, cockpit adjustment */
        bool adjustable = false;
        for (int id = 0;; ++id) {
            const std::string sec = "TURBO_" + std::to_string(id);
            if (!ini.hasSection(sec)) break;
            if (id >= PD_MAX_TURBOS) throw Error("more than PD_MAX_TURBOS turbochargers");
            PdTurbo& U = P.turbo[id]; memset(&U, 0, sizeof(U));
            U.lagDN = (1.0f - ini.getFloat(sec, "LAG_DN")) * 1.333333f * 333.3333f;
            U.lagUP = (1.0f - ini.getFloat(sec, "LAG_UP")) * 1.333333f * 333.3333f;
            U.maxBoost = ini.getFloat(sec, "MAX_BOOST"); U.wastegate = ini.getFloat(sec, "WASTEGATE");
            U.rpmRef = ini.getFloat(sec, "REFERENCE_RPM"); U.gamma = ini.getFloat(sec, "GAMMA");
            U.isAdjustable = ini.getInt(sec, "COCKPIT_ADJUSTABLE") != 0; U.userSetting = 1.0f;      /* Turbo::userSetting = 1 (Turbo.h) */
            if (U.isAdjustable) adjustable = true;
            P.nTurbos = id + 1;
        }
        if (adjustable) { const float boost = ini.getFloat("ENGINE_DATA", "DEFAULT_TURBO_ADJUSTMENT"); for (int i = 0; i < P.nTurbos; ++i) P.turbo[i].userSetting = P.turbo[i].isAdjustable ? boost : 1.0f; }
        for (int i = 0; i < P.nTurbos; ++i) if (file_exists(dataPath + "ctrl_turbo" + std::to_string(i) + ".ini") || file_exists(dataPath + "ctrl_wastegate" + std::to_string(i) + ".ini")) throw Error("turbo dynamic controllers are not supported yet");
    }
    if (ini.hasSection("OVERLAP")) { E.overlapFreq = ini.getFloat("OVERLAP", "FREQUENCY"); E.overlapGain = ini.getFloat("OVERLAP", "GAIN"); E.overlapIdealRPM = ini.getFloat("OVERLAP", "IDEAL_RPM"); }
    load_curve(dataPath + "throttle.lut", E.throttleResponseCurve);
    if (ini.hasSection("DAMAGE")) {
        E.rpmDamageThreshold = ini.getFloat("DAMAGE", "RPM_THRESHOLD"); E.rpmDamageK = ini.getFloat("DAMAGE", "RPM_DAMAGE_K");
        if (P.nTurbos > 0) { E.turboBoostDamageThreshold = ini.getFloat("DAMAGE", "TURBO_BOOST_THRESHOLD"); E.turboBoostDamageK = ini.getFloat("DAMAGE", "TURBO_DAMAGE_K"); }
    }
    if (ini.hasSection("BOV")) E.bovThreshold = ini.getFloat("BOV", "PRESSURE_THRESHOLD");
    if (ini.hasSection("THROTTLE_RESPONSE")) { E.throttleResponseCurveMaxRef = ini.getFloat("THROTTLE_RESPONSE", "RPM_REFERENCE"); E.throttleResponseCurveMax = ini.getCurve("THROTTLE_RESPONSE", "LUT"); }
    /* precalculatePowerAndTorque */
    const float fMaxRef = E.powerCurve.n ? E.powerCurve.ref[E.powerCurve.n - 1] : 0.0f;
    float maxTorqueNM = 0, maxPowerW = 0;
    for (float rpm = 0; rpm <= fMaxRef; rpm += 50.0f) {
        const float tq = curve_value(E.powerCurve, rpm);
        if (tq > maxTorqueNM) { E.maxTorqueRPM = rpm; maxTorqueNM = tq; }
        const float pw = rpm * tq * 0.1047f;
        if (pw > maxPowerW) { E.maxPowerRPM = rpm; maxPowerW = pw; }
    }
}

static void load_drivetrain(const std::string& dataPath, const PdCarParams& P, PdDrivetrain& D, PdAssists& A) {
    memset(&D, 0, sizeof(D)); memset(&A, 0, sizeof(A));
    Ini ini(dataPath + "drivetrain.ini");
    if (!ini.ready) throw Error("cannot read drivetrain.ini");
    const std::string trac = ini.getString("TRACTION", "TYPE");
    if (trac == "AWD" || trac == "AWD2") throw Error("AWD drivetrains are not implemented (reference: TODO_NOT_IMPLEMENTED_FATAL, Drivetrain.cpp:36-40)");
    D.damageRpmWindow = ini.getFloat("DAMAGE", "RPM_WINDOW_K");
    D.nGears = 0;
    D.gears[D.nGears++] = ini.getFloat("GEARS", "GEAR_R");
    D.gears[D.nGears++] = 0.0f;
    const int n = ini.getInt("GEARS", "COUNT");
    if (n + 2 > PD_MAX_GEARS) throw Error("too many gears");
    for (int i = 1; i <= n; ++i) D.gears[D.nGears++] = ini.getFloat("GEARS", "GEAR_" + std::to_string(i));
    D.finalRatio = ini.getFloat("GEARS", "FINAL");
    D.diffPowerRamp = ini.getFloat("DIFFERENTIAL", "POWER"); D.diffCoastRamp = ini.getFloat("DIFFERENTIAL", "COAST"); D.diffPreLoad = ini.getFloat("DIFFERENTIAL", "PRELOAD");
    D.diffType = (D.diffPowerRamp >= 1.0f && D.diffCoastRamp >= 1.0f) ? 1 : 0;
    if (trac == "RWD") { D.tractionType = 0; D.shaftInertiaL = P.tyre[2].angularInertia; D.shaftInertiaR = P.tyre[3].angularInertia; }
    else if (trac == "FWD") { D.tractionType = 1; D.shaftInertiaL = P.tyre[0].angularInertia; D.shaftInertiaR = P.tyre[1].angularInertia; }
    else throw Error("unknown TRACTION TYPE '" + trac + "'");
    D.gearUpTime = ini.getFloat("GEARBOX", "CHANGE_UP_TIME") * 0.001f; if (D.gearUpTime == 0.0f) D.gearUpTime = 0.1f;
    D.gearDnTime = ini.getFloat("GEARBOX", "CHANGE_DN_TIME") * 0.001f; if (D.gearDnTime == 0.0f) D.gearDnTime = 0.15f;
    D.autoCutOffTime = ini.getFloat("GEARBOX", "AUTO_CUTOFF_TIME") * 0.001f;
    D.isShifterSupported = ini.getInt("GEARBOX", "SUPPORTS_SHIFTER") != 0;
    double win = ini.getFloat("GEARBOX", "VALID_SHIFT_RPM_WINDOW"); if (win == 0.0) win = 500.0;
    D.orgRpmWindow = win;
    D.controlsWindowGain = ini.getFloat("GEARBOX", "CONTROLS_WINDOW_GAIN");
    D.clutchInertia = 1.0; D.driveInertia = 0.01f;
    const float gi = ini.getFloat("GEARBOX", "INERTIA"); if (gi != 0.0f) { D.clutchInertia = gi; D.driveInertia = gi; }
    D.clutchMaxTorque = ini.getFloat("CLUTCH", "MAX_TORQUE"); if (D.clutchMaxTorque == 0.0) D.clutchMaxTorque = 450.0;
    if (file_exists(dataPath + "ctrl_single_lock.ini")) throw Error("ctrl_single_lock.ini (dynamic diff controller) is not supported yet");
    /* AutoClutch */
    A.acRpmMin = 1500; A.acRpmMax = 2500; A.acClutchSpeed = 1.0f;
    const std::string up = ini.getString("AUTOCLUTCH", "UPSHIFT_PROFILE"), dn = ini.getString("AUTOCLUTCH", "DOWNSHIFT_PROFILE");
    A.acUseAutoOnChange = ini.getInt("AUTOCLUTCH", "USE_ON_CHANGES") != 0;
    auto prof = [&](const std::string& name, PdCurve& c) {
        if (name != "NONE" && ini.hasSection(name)) {
            const float p0 = ini.getFloat(name, "POINT_0"), p1 = ini.getFloat(name, "POINT_1"), p2 = ini.getFloat(name, "POINT_2");
            curve_add(c, 0.0f, 1.0f); curve_add(c, p0 * 0.001f, 0.0f); curve_add(c, p1 * 0.001f, 0.0f); curve_add(c, p2 * 0.001f, 1.0f);
        }
    };
    prof(up, A.upshiftProfile); prof(dn, A.downshiftProfile);
    A.acRpmMin = ini.getFloat("AUTOCLUTCH", "MIN_RPM"); A.acRpmMax = ini.getFloat("AUTOCLUTCH", "MAX_RPM");
    if (A.acRpmMin == 0.0f || A.acRpmMax == 0.0f) { A.acRpmMin = 1500.0f; A.acRpmMax = 2500.0f; }
    /* AutoBlip */
    {
        const float lvl = ini.getFloat("AUTOBLIP", "LEVEL"), p0 = ini.getFloat("AUTOBLIP", "POINT_0"), p1 = ini.getFloat("AUTOBLIP", "POINT_1"), p2 = ini.getFloat("AUTOBLIP", "POINT_2");
        curve_add(A.blipProfile, 0.0f, 0.0f); curve_add(A.blipProfile, p0, lvl); curve_add(A.blipProfile, p1, lvl); curve_add(A.blipProfile, p2, 0.0f);
        A.blipPerformTime = A.blipProfile.ref[A.blipProfile.n - 1];
        A.blipIsElectronic = ini.getInt("AUTOBLIP", "ELECTRONIC") != 0;
    }
    /* AutoShifter */
    A.asChangeUpRpm = 0; A.asChangeDnRpm = 4000; A.asSlipThreshold = 0.8f; A.asGasCutoffTime = 0.5f;
    if (ini.hasSection("AUTO_SHIFTER")) {
        A.asChangeUpRpm = ini.getInt("AUTO_SHIFTER", "UP"); A.asChangeDnRpm = ini.getInt("AUTO_SHIFTER", "DOWN");
        A.asSlipThreshold = ini.getFloat("AUTO_SHIFTER", "SLIP_THRESHOLD"); A.asGasCutoffTime = ini.getFloat("AUTO_SHIFTER", "GAS_CUTOFF_TIME");
    }
    if (!A.asChangeUpRpm) { /* AutoShifter.cpp:38-53 resolves these lazily on the first active step */
        const float lim = (float)(int)(P.engine.limiter * P.engine.limiterMultiplier);
        const float mx = lim >= P.engine.maxPowerRPM ? P.engine.maxPowerRPM : lim;
        A.asChangeUpRpm = (int)(mx * 0.98f); A.asChangeDnRpm = (int)(P.engine.maxTorqueRPM * 1.1f);
    }
}

static void load_aero(const std::string& dataPath, PdCarParams& P) {
    Ini ini(dataPath + "aero.ini");
    if (!ini.ready) throw Error("cannot read aero.ini");
    const int iVer = ini.getInt("HEADER", "VERSION");
    P.nWings = 0;
    for (int pass = 0; pass < 2; ++pass)
        for (int id = 0;; ++id) {
            const std::string sec = std::string(pass ? "FIN_" : "WING_") + std::to_string(id);
            if (!ini.hasSection(sec)) break;
            if (P.nWings >= PD_MAX_WINGS) throw Error("more than PD_MAX_WINGS wings");
            PdWing& W = P.wing[P.nWings++]; memset(&W, 0, sizeof(W));
            W.isVertical = pass; W.angleMult = 1.0f;
            const float chord = ini.getFloat(sec, "CHORD"), span = ini.getFloat(sec, "SPAN");
            W.area = chord * span; ini.getFloat3(sec, "POSITION", W.position);
            load_curve(dataPath + ini.getString(sec, "LUT_AOA_CL"), W.lutAOA_CL);
            load_curve(dataPath + ini.getString(sec, "LUT_AOA_CD"), W.lutAOA_CD);
            if (file_exists(dataPath + ini.getString(sec, "LUT_GH_CL")) || file_exists(dataPath + ini.getString(sec, "LUT_GH_CD")))
                throw Error("wing ground-height LUTs are not supported yet");
            W.cdGain = ini.getFloat(sec, "CD_GAIN"); W.clGain = ini.getFloat(sec, "CL_GAIN"); W.angle = ini.getFloat(sec, "ANGLE");
            if (iVer >= 3) W.yawGain = ini.getFloat(sec, "YAW_CL_GAIN");
        }
    if (P.nWings == 0) throw Error("aero.ini without wings ([DATA] drag/lift model) is not supported yet");
    if (ini.hasSection("DYNAMIC_CONTROLLER_0")) throw Error("wing dynamic controllers are not supported yet");
}

void load_car(const std::string& basePathIn, const std::string& model, CarModel& out) {
    std::string basePath = basePathIn;
    if (!basePath.empty() && basePath.back() != '/') basePath += "/";
    const std::string dataPath = basePath + "content/cars/" + model + "/data/";
    out.dataPath = dataPath; out.setupVars.clear();
    PdCarParams& P = out.P; memset(&P, 0, sizeof(P));
    /* ---- Simulator::init ---- */
    P.roadTemperature = 20.0f; P.ambientTemperature = 20.0f; P.fuelConsumptionRate = 0.0f; P.tyreConsumptionRate = 0.0f; P.mechanicalDamageRate = 1.0f;
    {
        Ini sim(basePath + "cfg/sim.ini");
        if (sim.ready) { sim.tryGetFloat("ENVIRONMENT", "ROAD_TEMP", P.roadTemperature); sim.tryGetFloat("ENVIRONMENT", "AMBIENT_TEMP", P.ambientTemperature); }
        /* Car::initProbes / initLookAhead */
        P.lookAheadCount = 5; P.lookAheadStep = 10.0f; P.nProbes = 0;
        if (sim.ready) {
            for (int id = 1; id <= PD_MAX_PROBES; ++id) {
                const std::string sec = "CAR_PROBE_" + std::to_string(id);
                if (!sim.hasSection(sec)) break;
                const float yaw = sim.getFloat(sec, "YAW"), length = sim.getFloat(sec, "LENGTH");
                /* vec3f(0,0,1).rotateAxisAngle((0,1,0), yaw*DEG2RAD) = (0,0,1) * createFromAxisAngle -> third row */
                const float a = yaw * (float)0.01745329251994329576923690768489, s = sinf(a), c = cosf(a), o = 1.0f - c;
                const float m31 = (1.0f * s) + (0.0f * 0.0f) * o, m32 = (0.0f * 1.0f) * o - (0.0f * s), m33 = ((0.0f * 0.0f) * o) + c;
                P.probeDir[P.nProbes][0] = 0.0f + (0.0f * 0.0f + 0.0f * 0.0f + 1.0f * m31); P.probeDir[P.nProbes][1] = 0.0f + (0.0f + 0.0f + 1.0f * m32); P.probeDir[P.nProbes][2] = 0.0f + (0.0f + 0.0f + 1.0f * m33);
                P.probeLength[P.nProbes] = length; P.nProbes++;
            }
            sim.tryGetInt("CAR_LOOK_AHEAD", "COUNT", P.lookAheadCount); sim.tryGetFloat("CAR_LOOK_AHEAD", "STEP", P.lookAheadStep);
        }
        if (P.lookAheadCount > PD_LOOKAHEAD) throw Error("CAR_LOOK_AHEAD COUNT > 5");
    }
    P.airDensity = 1.2922f - (P.ambientTemperature * 0.0041f);
    P.gravityY = -9.80665f; P.worldERP = 0.3f; P.worldCFM = 1.0e-7f;
    /* ---- Car::initCarData ---- */
    Ini car(dataPath + "car.ini");
    if (!car.ready) throw Error("cannot read " + dataPath + "car.ini");
    P.mass = car.getFloat("BASIC", "TOTALMASS");
    /* the reference writes the three explicit moments into slots [0], [4], [8] of ODE's 3 x 4 inertia matrix (RigidBodyODE.cpp:80-89, its own "TODO: why not
       [0] [5] [10]?"): the tensor it hands to dBodySetMass is singular, such a car cannot be simulated by the reference itself */
    if (car.hasSection("EXPLICIT_INERTIA")) throw Error("EXPLICIT_INERTIA cars are not supported (the reference builds a singular inertia tensor for them, RigidBodyODE.cpp:80-89)");
    float bodyInertia[3]; car.getFloat3("BASIC", "INERTIA", bodyInertia);
    P.fuelKG = 0.74f; if (car.hasSection("FUEL_EXT")) P.fuelKG = car.getFloat("FUEL_EXT", "KG_PER_LITER");
    P.steerLock = car.getFloat("CONTROLS", "STEER_LOCK"); P.steerRatio = car.getFloat("CONTROLS", "STEER_RATIO");
    P.steerLinearRatio = car.getFloat("CONTROLS", "LINEAR_STEER_ROD_RATIO"); if (P.steerLinearRatio == 0.0f) P.steerLinearRatio = 0.003f;
    P.fuelConsumptionK = car.getFloat("FUEL", "CONSUMPTION");
    double fuel = car.getFloat("FUEL", "FUEL"); P.maxFuel = car.getFloat("FUEL", "MAX_FUEL");
    if (P.maxFuel == 0.0f) P.maxFuel = 30.0f;
    if (fuel == 0.0f) fuel = 30.0f;
    P.requestedFuel = (float)fuel;
    car.getFloat3("FUELTANK", "POSITION", P.fuelTankPos);
    { /* ---- colliders: CarColliderManager::init (CarColliderManager.cpp:12-34) + Car::loadColliderBlob (Car.cpp:318-379) ---- */
        Ini col(dataPath + "colliders.ini");
        if (!col.ready) throw Error("cannot read " + dataPath + "colliders.ini");
        if (col.hasSection("COLLIDER_0")) { P.hasBoxCollider = 1; col.getFloat3("COLLIDER_0", "CENTRE", P.boxCentre); col.getFloat3("COLLIDER_0", "SIZE", P.boxSize); }
        if (col.hasSection("COLLIDER_1")) throw Error("cars with more than one box collider are not supported yet");
        float goff[3]; car.getFloat3("BASIC", "GRAPHICS_OFFSET", goff);
        /* geom offset = Car::getGraphicsOffsetMatrix() at construction (Car.cpp:1407-1424): translation GRAPHICS_OFFSET, rotation =
           mat44f::createFromAxisAngle((1,0,0), GRAPHICS_PITCH_ROTATION) (Core/Math.cpp:88-118), handed to ODE transposed
           (RigidBodyODE.cpp:300-312): local = r * v + offset */
        const float pitch = (float)(car.getFloat("BASIC", "GRAPHICS_PITCH_ROTATION") * (M_PI / 180.0f));
        float r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (pitch != 0.0f) {
            const float sin_a = sinf(pitch), cos_a = cosf(pitch), om = 1.0f - cos_a;
            const float ax = 1.0f, ay = 0.0f, az = 0.0f;
            const float M11 = ((ax * ax) * om) + cos_a, M22 = ((ay * ay) * om) + cos_a, M33 = ((az * az) * om) + cos_a;
            const float M12 = (az * sin_a) + (ay * ax) * om, M23 = (ax * sin_a) + (az * ay) * om, M31 = (ay * sin_a) + (az * ax) * om;
            const float M13 = (az * ax) * om - (ay * sin_a), M21 = (ay * ax) * om - (az * sin_a), M32 = (az * ay) * om - (ax * sin_a);
            const float rr[9] = {M11, M21, M31, M12, M22, M32, M13, M23, M33};
            for (int k = 0; k < 9; ++k) r[k] = rr[k];
        }
        FILE* f = fopen((dataPath + "collider.bin").c_str(), "rb");
        if (!f) throw Error("cannot read " + dataPath + "collider.bin");
        uint32_t hdr[3] = {0, 0, 0};
        bool ok = fread(hdr, 4, 3, f) == 3 && hdr[1] > 0 && hdr[2] > 0 && hdr[1] <= PD_MAX_COLLIDER_VERTS && hdr[2] % 3 == 0 && hdr[2] / 3 <= PD_MAX_COLLIDER_TRIS;
        std::vector<float> v(ok ? hdr[1] * 3 : 0); std::vector<uint16_t> ix(ok ? hdr[2] : 0);
        ok = ok && fread(v.data(), 12, hdr[1], f) == hdr[1] && fread(ix.data(), 2, hdr[2], f) == hdr[2];
        fclose(f);
        if (!ok) throw Error("collider.bin: malformed or larger than PD_MAX_COLLIDER_VERTS / PD_MAX_COLLIDER_TRIS");
        P.nColliderVerts = (int32_t)hdr[1]; P.nColliderTris = (int32_t)(hdr[2] / 3);
        for (int k = 0; k < 3; ++k) { P.colliderMin[k] = 3.4e38f; P.colliderMax[k] = -3.4e38f; }
        for (int i = 0; i < P.nColliderVerts; ++i) for (int k = 0; k < 3; ++k) {
            const float x = (r[k * 3 + 0] * v[i * 3 + 0] + r[k * 3 + 1] * v[i * 3 + 1] + r[k * 3 + 2] * v[i * 3 + 2]) + goff[k];
            P.colliderVerts[i][k] = x; P.colliderMin[k] = std::min(P.colliderMin[k], x); P.colliderMax[k] = std::max(P.colliderMax[k], x);
        }
        for (int t = 0; t < P.nColliderTris; ++t) for (int k = 0; k < 3; ++k) {
            if (ix[t * 3 + k] >= hdr[1]) throw Error("collider.bin: index out of range");
            P.colliderTris[t][k] = (uint8_t)ix[t * 3 + k];
        }
        for (int t = 0; t < P.nColliderTris; ++t) for (int k = 0; k < 3; ++k) {
            const float a = P.colliderVerts[P.colliderTris[t][0]][k], b = P.colliderVerts[P.colliderTris[t][1]][k], c = P.colliderVerts[P.colliderTris[t][2]][k];
            P.colliderTriBounds[t][k] = std::min(a, std::min(b, c)) - 1e-4f; P.colliderTriBounds[t][3 + k] = std::max(a, std::max(b, c)) + 1e-4f;
        }
        for (int t = 0; t < P.nColliderTris; ++t) {   /* bounding sphere about the box centre */
            float c[3], r2 = 0.0f;
            for (int k = 0; k < 3; ++k) c[k] = 0.5f * (P.colliderTriBounds[t][k] + P.colliderTriBounds[t][3 + k]);
            for (int v = 0; v < 3; ++v) { float d2 = 0.0f; for (int k = 0; k < 3; ++k) { const float d = P.colliderVerts[P.colliderTris[t][v]][k] - c[k]; d2 += d * d; } r2 = std::max(r2, d2); }
            P.colliderTriSphere[t][0] = c[0]; P.colliderTriSphere[t][1] = c[1]; P.colliderTriSphere[t][2] = c[2]; P.colliderTriSphere[t][3] = sqrtf(r2) * 1.0001f + 1e-4f;
        }
    }
    P.framesToSleep = 50;
    P.waterTmass = 20.0f; P.waterCoolSpeedK = 0.002f; P.waterCoolFactor = 0.2f; P.waterHeatFactor = 1.0f;
    /* ---- bodies at construction: chassis at the origin, identity rotation (Car.cpp:38-51) ---- */
    HBody chassis, tank;
    tank.setPos(mk(P.fuelTankPos));
    { /* dJointSetFixed(tank, chassis) */
        const V ofs = V{tank.pos[0] - chassis.pos[0], tank.pos[1] - chassis.pos[1], tank.pos[2] - chassis.pos[2]};
        put(P.tankOffset, tank.vecToLocal(ofs)); qmul1(P.tankQrel, tank.q, chassis.q);
    }
    /* ---- suspensions ---- */
    Ini susp(dataPath + "suspensions.ini");
    if (!susp.ready) throw Error("cannot read suspensions.ini");
    const std::string frontType = susp.getString("FRONT", "TYPE"), rearType = susp.getString("REAR", "TYPE");
    if ((frontType != "STRUT" && frontType != "DWB" && frontType != "ML") || (rearType != "AXLE" && rearType != "DWB" && rearType != "ML"))
        throw Error("unknown suspension TYPE (STRUT / DWB / ML front, AXLE / DWB / ML rear)");
    /* a multilink corner (SuspensionML) is a hub on five distance joints like a double-wishbone one: same kernels, PdDW::multilink = 1 */
    const bool frontML = frontType == "ML", rearML = rearType == "ML";
    const bool frontDW = frontType == "DWB" || frontML, rearDW = rearType == "DWB" || rearML;
    P.topology = (frontDW ? 2 : 0) + (rearDW ? 1 : 0);
    if (P.topology == PD_TOPO_DW_AXLE) throw Error("DWB front with a rigid rear axle: no kernel instance is built for this pair");
    float hubMass[2];
    if (frontML) { load_ml(susp, 0, chassis, P.dw[0]); load_ml(susp, 1, chassis, P.dw[1]); }
    else if (frontDW) { load_dw(susp, 0, chassis, P.dw[0]); load_dw(susp, 1, chassis, P.dw[1]); load_heave(susp, true, P.heave[0]); }
    else { load_strut(susp, 0, chassis, P.strut[0], hubMass[0]); load_strut(susp, 1, chassis, P.strut[1], hubMass[1]); }
    if (rearML) { load_ml(susp, 2, chassis, P.dw[2]); load_ml(susp, 3, chassis, P.dw[3]); }
    else if (rearDW) { load_dw(susp, 2, chassis, P.dw[2]); load_dw(susp, 3, chassis, P.dw[3]); load_heave(susp, false, P.heave[1]); }
    else load_axle(susp, chassis, P.axle);
    P.arbK[0] = susp.getFloat("ARB", "FRONT"); P.arbK[1] = susp.getFloat("ARB", "REAR");
    if (file_exists(dataPath + "ctrl_arb_front.ini") || file_exists(dataPath + "ctrl_arb_rear.ini")) throw Error("ARB dynamic controllers are not supported yet");
    /* ---- tyres ---- */
    for (int w = 0; w < 4; ++w) load_tyre(dataPath, w, P.ambientTemperature, P.tyre[w]);
    /* Car::getBaseCarHeight */
    {
        const float t0 = fabsf((frontDW ? P.dw[0].refPoint[1] : P.strut[0].refPoint[1]) - P.tyre[0].rimRadius), t2 = fabsf((rearDW ? P.dw[2].refPoint[1] : P.axle.axleBasePos[1]) - P.tyre[2].rimRadius);
        P.baseCarHeight = t0 > t2 ? t0 : t2;
    }
    /* ---- components ---- */
    load_aero(dataPath, P);
    { /* BrakeSystem::init */
        Ini b(dataPath + "brakes.ini"); if (!b.ready) throw Error("cannot read brakes.ini");
        P.brakes.brakePower = b.getFloat("DATA", "MAX_TORQUE"); P.brakes.frontBias = b.getFloat("DATA", "FRONT_SHARE");
        P.brakes.handBrakeTorque = b.getFloat("DATA", "HANDBRAKE_TORQUE"); P.brakes.brakePowerMultiplier = 1.0f;
        P.brakes.biasMin = 0; P.brakes.biasMax = 1.0f;
        if (file_exists(dataPath + "ctrl_ebb.ini") || file_exists(dataPath + "steer_brake_controller.ini"))
            throw Error("brake dynamic controllers (ctrl_ebb.ini, steer_brake_controller.ini) are not supported yet (SURVEY.md N4)");
        if (b.hasSection("EBB")) { P.brakes.ebbInternal = 1; const float m = b.getFloat("EBB", "FRONT_SHARE_MULTIPLIER"); P.brakes.ebbFrontMultiplier = m > 1.1f ? m : 1.1f; }
        if (b.hasSection("TEMPS_FRONT") && b.hasSection("TEMPS_REAR")) { /* BrakeSystem.cpp:41-54 */
            P.brakes.hasTemps = 1;
            for (int id = 0; id < 4; ++id) {
                const std::string sec = id < 2 ? "TEMPS_FRONT" : "TEMPS_REAR";
                PdBrakeDisc& D = P.brakes.disc[id];
                D.perfCurve = b.getCurve(sec, "PERF_CURVE"); D.torqueK = b.getFloat(sec, "TORQUE_K");
                D.coolTransfer = b.getFloat(sec, "COOL_TRANSFER"); D.coolSpeedFactor = b.getFloat(sec, "COOL_SPEED_FACTOR");
            }
        }
        Ini s(dataPath + "setup.ini");
        if (s.ready && s.hasSection("FRONT_BIAS")) { P.brakes.biasMin = s.getFloat("FRONT_BIAS", "MIN") * 0.01f; P.brakes.biasMax = s.getFloat("FRONT_BIAS", "MAX") * 0.01f; }
    }
    load_engine(dataPath, P.engine, P);
    load_drivetrain(dataPath, P, P.drivetrain, P.assists);
    /* tyres[].driven */
    for (int w = 0; w < 4; ++w) P.tyre[w].driven = (P.drivetrain.tractionType == 0) ? (w >= 2) : (w < 2);
    /* scoring defaults */
    {
        const float def[PD_NUM_SCORING_VARS] = {10.0f, 5.0f, 200.0f, 300.0f, 0.75f, 0.51f, 3.0f, 2.0f, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        memcpy(P.scoring, def, sizeof(def));
    }
    /* ---- Car::updateBodyMass (Car.cpp:589-620) at init ---- */
    {
        /* Car::calcBodyMass: fSuspMass += suspensions[i]->getMass() for i = 0..3, in that order */
        float suspMass = 0.0f;
        for (int w = 0; w < 4; ++w) suspMass += (w < 2) ? (frontDW ? P.dw[w].hubMass : P.strut[w].hubMass) : (rearDW ? P.dw[w].hubMass : (P.axle.axleMass * 0.5f));
        P.chassisMass = (P.mass - suspMass) + 0.0f;
        box_inertia(P.chassisMass, bodyInertia[0], bodyInertia[1], bodyInertia[2], P.chassisInertia);
        const float fuelMass = (P.fuelKG * (float)fuel) > 0.1f ? (P.fuelKG * (float)fuel) : 0.1f;
        P.tankMass = fuelMass; box_inertia(fuelMass, 0.5f, 0.5f, 0.5f, P.tankInertia);
    }
    /* ---- SetupManager (only the variables the kernels read) ---- */
    auto addF = [&](const char* name, double mult, float* p) { SetupVar v; v.name = name; v.fvalue = p; v.mult = (float)mult; out.setupVars.push_back(v); };
    auto addD = [&](const char* name, double mult, double* p) { SetupVar v; v.name = name; v.dvalue = p; v.mult = (float)mult; out.setupVars.push_back(v); };
    addF("FRONT_BIAS", 0.01, &P.brakes.frontBias); addF("BRAKE_POWER_MULT", 0.01, &P.brakes.brakePowerMultiplier);
    addD("DIFF_POWER", 0.01, &P.drivetrain.diffPowerRamp); addD("DIFF_COAST", 0.01, &P.drivetrain.diffCoastRamp); addD("DIFF_PRELOAD", 1.0, &P.drivetrain.diffPreLoad);
    addD("FINAL_RATIO", 1.0, &P.drivetrain.finalRatio);
    addF("ARB_FRONT", 1.0, &P.arbK[0]); addF("ARB_REAR", 1.0, &P.arbK[1]);
    addF("ENGINE_LIMITER", 0.01, &P.engine.limiterMultiplier);
    for (int i = 0; i < P.nTurbos; ++i) { /* SetupManager.cpp:94-107: TURBO_n -> Turbo::userSetting, tunable 0..1 in steps of 0.1 */
        static char nm[PD_MAX_TURBOS][16]; snprintf(nm[i], sizeof(nm[i]), "TURBO_%d", i); addF(nm[i], 0.01, &P.turbo[i].userSetting);
        SetupVar& v = out.setupVars.back(); v.minV = 0.0f; v.maxV = 1.0f; v.step = 0.1f; v.tunable = true;
    }
    for (int i = 0; i < P.drivetrain.nGears && i < PD_MAX_GEARS; ++i) { static char nm[PD_MAX_GEARS][32]; snprintf(nm[i], sizeof(nm[i]), "INTERNAL_GEAR_%d", i); addD(nm[i], 1.0, &P.drivetrain.gears[i]); }
    {   /* front struts: the per-wheel variables of SetupManager.cpp:62-83 */
        static const char* kSide[2] = {"LF", "RF"}; static const double kSign[2] = {-1, 1};
        static char nm[2][11][40];
        for (int w = 0; w < 2 && !frontDW; ++w) {
            PdStrut& S = P.strut[w]; int q = 0;
            auto reg = [&](const char* base, double mult, float* ptr) { snprintf(nm[w][q], sizeof(nm[w][q]), "%s_%s", base, kSide[w]); addF(nm[w][q], mult, ptr); ++q; };
            reg("DAMP_FAST_BUMP", 1.0, &S.damper.bumpFast); reg("DAMP_BUMP", 1.0, &S.damper.bumpSlow);
            reg("DAMP_FAST_REBOUND", 1.0, &S.damper.reboundFast); reg("DAMP_REBOUND", 1.0, &S.damper.reboundSlow);
            reg("BUMP_STOP_RATE", 1000.0, &S.bumpStopRate); reg("SPRING_RATE", 1000.0, &S.k); reg("PROGRESSIVE_SPRING_RATE", 1000.0, &S.progressiveK);
            reg("ROD_LENGTH", 0.0001, &S.rodLength); reg("CAMBER", 0.0017453292 * kSign[w], &S.staticCamber);
            reg("TOE_OUT", 0.00001, &S.toeOutLinear); reg("PACKER_RANGE", 0.001, &S.packerRange);
        }
    }
    {   /* double-wishbone corners: the same per-wheel variables (SetupManager.cpp:62-83), all four wheels where the axle is DWB */
        static const char* kSide4[4] = {"LF", "RF", "LR", "RR"}; static const double kSign4[4] = {-1, 1, -1, 1};
        static char nm4[4][11][40];
        for (int w = 0; w < 4; ++w) {
            if (!(w < 2 ? frontDW : rearDW)) continue;
            PdDW& S = P.dw[w]; int q = 0;
            auto reg = [&](const char* base, double mult, float* ptr) { snprintf(nm4[w][q], sizeof(nm4[w][q]), "%s_%s", base, kSide4[w]); addF(nm4[w][q], mult, ptr); ++q; };
            reg("DAMP_FAST_BUMP", 1.0, &S.damper.bumpFast); reg("DAMP_BUMP", 1.0, &S.damper.bumpSlow);
            reg("DAMP_FAST_REBOUND", 1.0, &S.damper.reboundFast); reg("DAMP_REBOUND", 1.0, &S.damper.reboundSlow);
            reg("BUMP_STOP_RATE", 1000.0, &S.bumpStopRate); reg("SPRING_RATE", 1000.0, &S.k); reg("PROGRESSIVE_SPRING_RATE", 1000.0, &S.progressiveK);
            reg("ROD_LENGTH", 0.0001, &S.rodLength); reg("CAMBER", 0.0017453292 * kSign4[w], &S.staticCamber);
            reg("TOE_OUT", 0.00001, &S.toeOutLinear); reg("PACKER_RANGE", 0.001, &S.packerRange);
        }
    }
    for (int i = 0; i < P.nWings && i < PD_MAX_WINGS; ++i) { static char nm[PD_MAX_WINGS][16]; snprintf(nm[i], sizeof(nm[i]), "WING_%d", i); addF(nm[i], 1.0, &P.wing[i].angle); }
    /* PRESSURE_xx tune status.pressureStatic (per-env state): handled by the batch through pressureStaticDefault */
    addF("PRESSURE_LF", 1.0, &P.tyre[0].pressureStaticDefault); addF("PRESSURE_RF", 1.0, &P.tyre[1].pressureStaticDefault);
    addF("PRESSURE_LR", 1.0, &P.tyre[2].pressureStaticDefault); addF("PRESSURE_RR", 1.0, &P.tyre[3].pressureStaticDefault);
    Ini setup(dataPath + "setup.ini");
    if (setup.ready) {
        std::string ratiosFile;
        if (setup.tryGetString("FINAL_GEAR_RATIO", "RATIOS", ratiosFile)) {
            /* SetupGearRatio::load */
            std::vector<float> vals;
            FILE* f = fopen((dataPath + ratiosFile).c_str(), "rb");
            if (f) {
                char line[256];
                while (fgets(line, sizeof(line), f)) { std::string l(line); while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back(); if (l.empty()) continue; auto kv = split(l, "|"); if (kv.size() == 2) vals.push_back(stof_ref(kv[1])); }
                fclose(f);
            }
            std::sort(vals.begin(), vals.end());
            if (!vals.empty()) for (auto& v : out.setupVars) if (v.name == "FINAL_RATIO") { v.spinnerType = 4; v.spinnerValues = vals; v.mult = 1; v.minV = vals.front(); v.maxV = vals.back(); v.step = 0.01f; v.tunable = true; }
        }
        for (auto& v : out.setupVars) {
            if (!setup.hasSection(v.name)) continue;
            const bool tunable = setup.tryGetFloat(v.name, "MIN", v.minV) && setup.tryGetFloat(v.name, "MAX", v.maxV) && setup.tryGetFloat(v.name, "STEP", v.step);
            int clicks = 0;
            if (setup.tryGetInt(v.name, "SHOW_CLICKS", clicks) && clicks >= 0 && clicks <= 4) v.spinnerType = clicks;
            if (tunable) v.tunable = true;
        }
    }
    for (auto& v : out.setupVars) {
        if (v.tunable) {
            if (v.minV >= v.maxV || fabs(v.maxV - v.minV) < 0.01) { v.minV = 0; v.maxV = 0; v.tunable = false; }
            if (v.step < 0.0f) { v.step = 0.0f; v.tunable = false; }
        }
        if (!v.tunable) v.spinnerType = 3;
    }
    if (P.tyreConsumptionRate > 0.0f) throw Error("tyre consumption (grain / blister / wear) is not supported yet");
}

static inline float trunc_f(float x) { return (float)(int)x; }

/* SetupVar::setTune = getSpinner + clamp + setValue (SetupManager.cpp:283-399) */
/* variables the reference registers (SetupManager.cpp:18-135) that point at state this build keeps per AXLE, not per wheel, or
 * at components the demo-car kernels do not have (AWD differentials, turbos): setting them must fail loudly, not silently */
static bool known_but_unsupported(const std::string& name) {
    static const char* kPrefix[] = {"FRONT_DIFF_", "REAR_DIFF_", "CENTER_DIFF_", "AWD_FRONT_TORQUE_DISTRIBUTION", "TURBO_", nullptr};
    for (int i = 0; kPrefix[i]; ++i) if (name.compare(0, strlen(kPrefix[i]), kPrefix[i]) == 0) return true;
    static const char* kWheel[] = {"DAMP_FAST_BUMP_", "DAMP_BUMP_", "DAMP_FAST_REBOUND_", "DAMP_REBOUND_", "BUMP_STOP_RATE_", "SPRING_RATE_", "PROGRESSIVE_SPRING_RATE_",
                                   "ROD_LENGTH_", "CAMBER_", "TOE_OUT_", "PACKER_RANGE_", nullptr};
    for (int i = 0; kWheel[i]; ++i) {
        const size_t n = strlen(kWheel[i]);
        if (name.size() == n + 2 && name.compare(0, n, kWheel[i]) == 0 && (name.compare(n, 2, "LR") == 0 || name.compare(n, 2, "RR") == 0)) return true;
    }
    return false;
}
/* SetupVar::setRaw: the value as is, no spinner clamp / multiplier (SetupManager.cpp:276-281) */
void CarModel::setRawTune(const std::string& name, float v) {
    for (auto& var : setupVars) {
        if (var.name != name) continue;
        if (var.fvalue) *var.fvalue = v; else if (var.dvalue) *var.dvalue = v;
        return;
    }
    if (known_but_unsupported(name)) throw Error("setup variable '" + name + "' exists in the reference but is not supported by this build (per-wheel rear-axle / AWD / turbo tune)");
}
void CarModel::setTune(const std::string& name, float v) {
    bool registered = false;
    for (auto& var : setupVars) if (var.name == name) registered = true;
    if (!registered && known_but_unsupported(name)) throw Error("setup variable '" + name + "' exists in the reference but is not supported by this build (per-wheel rear-axle / AWD tune)");
    for (auto& var : setupVars) {
        if (var.name != name) continue;
        float smin = 0, smax = 0;
        switch (var.spinnerType) {
        case 0: smin = trunc_f(var.minV); smax = trunc_f(var.maxV); break;
        case 1: smin = trunc_f(var.minV / var.step); smax = trunc_f(var.maxV / var.step); break;
        case 2: smin = 0.0f; smax = trunc_f((var.maxV - var.minV) / var.step); break;
        default: smin = var.minV; smax = var.maxV; break;
        }
        const float value = v < smin ? smin : (v > smax ? smax : v);
        float raw = 0;
        switch (var.spinnerType) {
        case 0: raw = value * var.mult; break;
        case 1: raw = (value * var.step) * var.mult; break;
        case 2: raw = (value * var.step + var.minV) * var.mult; break;
        default: raw = value; break;
        }
        if (var.spinnerType == 4 && !var.spinnerValues.empty()) {
            int best = 0; float bd = 3.402823466e+38f;
            for (int i = 0; i < (int)var.spinnerValues.size(); ++i) { const float d = fabsf(var.spinnerValues[i] - raw); if (d < bd) { bd = d; best = i; } }
            raw = var.spinnerValues[best];
        }
        if (var.fvalue) *var.fvalue = raw; else if (var.dvalue) *var.dvalue = raw;
        return;
    }
    /* unknown names are ignored, as SetupManager::setTune does (getVar returns nullptr) */
}
void CarModel::setScoringVar(const std::string& name, float value) {
    for (int i = 0; i < PD_NUM_SCORING_VARS; ++i) if (name == kScoringVarNames[i]) { P.scoring[i] = value; return; }
    throw Error("unknown scoring variable '" + name + "'");
}
float CarModel::getScoringVar(const std::string& name) const {
    for (int i = 0; i < PD_NUM_SCORING_VARS; ++i) if (name == kScoringVarNames[i]) return P.scoring[i];
    return 0.0f;
}

} // namespace pdh
