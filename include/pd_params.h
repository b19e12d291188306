/*
 * pd_params.h -- read-only parameter block of one car model (+ simulator constants), as the kernels
 * consume it.  Plain C POD, filled once on the host by the loader (projectd_core_b200/csrc/host/
 * car_loader.cpp) from the reference's own data formats (content/cars/<model>/data/ *.ini, *.lut;
 * cfg/sim.ini) following the reference's init code, then copied to the device once.
 *
 * The oracle harness fills the same struct from the reference's live objects so that the loader can be
 * checked field by field (tests/test_loader_parity.py).
 *
 * Reference anchors (src/ProjectD/...): Car/Car.cpp:31-316 (Car::init, initCarData, initProbes),
 * Car/SuspensionStrut.cpp:20-186, Car/SuspensionAxle.cpp:16-113, Car/Tyre.cpp:48-392,
 * Car/Engine.cpp:17-168, Car/Drivetrain.cpp:19-152, Car/AutoClutch.cpp:28-89, Car/AutoBlip.cpp:17-49,
 * Car/AutoShifter.cpp:15-29, Car/BrakeSystem.cpp:15-71, Car/AeroMap.cpp:16-81, Car/Wing.cpp:20-69,
 * Car/ScoringSystem.cpp:50-73, Sim/Simulator.cpp:22-90.
 */
#ifndef PD_PARAMS_H
#define PD_PARAMS_H

#include <stdint.h>
#include "pd_state.h"

#define PD_CURVE_MAX 24
#define PD_MAX_WINGS 6       /* bundled cars: 3 wings + up to 2 fins */
#define PD_MAX_GEARS 10
#define PD_AXLE_LINKS 5      /* SuspensionAxle.cpp:50 LINK_COUNT (demo car: 5) */
#define PD_NUM_SCORING_VARS 21

#ifdef __cplusplus
extern "C" {
#endif

/* Core/Curve.cpp:94-115 piece-wise linear LUT with end clamps */
typedef struct PdCurve {
    int32_t n;
    float ref[PD_CURVE_MAX];
    float val[PD_CURVE_MAX];
} PdCurve;

typedef struct PdDamper { /* Car/Damper.h:14-19 */
    float bumpSlow, reboundSlow, bumpFast, reboundFast, fastThresholdBump, fastThresholdRebound;
} PdDamper;

typedef struct PdDBall { /* dxJointDBall: body-local anchors + target distance */
    float anchor1[3];   /* on body 0 (chassis) */
    float anchor2[3];   /* on body 1 (hub / axle) */
    float distance;
} PdDBall;

typedef struct PdStrut { /* SuspensionStrut (front wheels of the demo car) */
    float refPoint[3];        /* dataRelToWheel.refPoint == basePosition (chassis frame) */
    float carStrut[3];        /* dataRelToBody.carStrut  (chassis frame) */
    float tyreStrut[3];       /* dataRelToWheel.tyreStrut (hub frame) */
    float baseCarSteer[3];    /* baseCarSteerPosition (chassis frame) */
    float tyreSteer[3];       /* dataRelToWheel.tyreSteer (hub frame) */
    float strutBaseLength, strutBodyLength;
    float rodLength, k, progressiveK, packerRange, bumpStopRate, bumpStopUp, bumpStopDn;
    float toeOutLinear, staticCamber, baseCFM;
    PdDamper damper;
    PdDBall link[3];          /* joints[0..2]: WB rear, WB front, steer rod (chassis <-> hub) */
    /* slider (strutBody, hub): axis in strutBody frame, offset in hub frame, qrel */
    float sliderAxis1[3], sliderOffset[3], sliderQrel[4];
    /* ball (chassis, strutBody): anchors */
    float ballAnchor1[3], ballAnchor2[3];
    float hubMass, hubInertia[3], strutMass, strutInertia[3];
} PdStrut;

typedef struct PdAxle { /* SuspensionAxle x2 sharing Car::rigidAxle (rear of the demo car) */
    float track, referenceY, attachRelativePos, axleBasePos[3], leafSpringKx;
    float rodLength, k, progressiveK, bumpStopUp, bumpStopDn, bumpStopRate, baseCFM;
    PdDamper damper;
    PdDBall link[PD_AXLE_LINKS];
    int32_t nLinks;
    float axleMass, axleInertia[3];
    float torqueReaction;     /* Car::axleTorqueReaction */
} PdAxle;

/* SuspensionDW (Car/SuspensionDW.cpp:18-195): one hub body held by five distance joints (upper wishbone rear / front, lower
 * wishbone rear / front, steering or toe rod), spring + damper along the chassis' up axis at the reference point */
#define PD_DW_LINKS 5
typedef struct PdDW {
    float refPoint[3];        /* dataRelToWheel.refPoint == basePosition (chassis frame) */
    float baseCarSteer[3];    /* baseCarSteerPosition (chassis frame) */
    float tyreSteer[3];       /* dataRelToWheel.tyreSteer (hub frame) */
    float rodLength, k, progressiveK, packerRange, bumpStopRate, bumpStopProgressive, bumpStopUp, bumpStopDn;
    float toeOutLinear, staticCamber, baseCFM;
    PdDamper damper;
    PdDBall link[PD_DW_LINKS]; /* joints[0..4]: top rear, top front, bottom rear, bottom front, steer rod (chassis <-> hub) */
    float hubMass, hubInertia[3];
    /* 1: SuspensionML (Car/SuspensionML.cpp:15-137) -- the same hub + five distance joints (JOINTn_CAR / JOINTn_TYRE), spring force applied
     * whatever its sign, plain packer, no bump stops, joints left at the world's ERP / CFM (its setERPCFM is empty) */
    int32_t multilink, pad0;
} PdDW;

/* HeaveSpring (Car/HeaveSpring.cpp:11-149): a third spring + damper between the chassis and the MEAN travel of the two hubs of a double-wishbone
 * axle, suspensions.ini [HEAVE_FRONT] / [HEAVE_REAR] */
typedef struct PdHeave {
    int32_t present, pad0;
    float bumpStopUp, bumpStopDn, rodLength, k, progressiveK, bumpStopRate, packerRange;
    PdDamper damper;
} PdHeave;

/* Turbo (Car/Turbo.h, Engine.cpp:69-94) */
#define PD_MAX_TURBOS 3
typedef struct PdTurbo {
    float lagDN, lagUP, maxBoost, wastegate, rpmRef, gamma, userSetting;
    int32_t isAdjustable;
} PdTurbo;

/* suspension topology = (front == DWB) * 2 + (rear == DWB); front otherwise STRUT, rear otherwise AXLE */
#define PD_TOPO_STRUT_AXLE 0
#define PD_TOPO_STRUT_DW   1
#define PD_TOPO_DW_AXLE    2
#define PD_TOPO_DW_DW      3

typedef struct PdTyre { /* TyreData + TyreModelData + SCTM + thermal patch data of the active compound */
    /* TyreData */
    float width, radius, rimRadius, k, d, angularInertia;
    float thermalFrictionK, thermalRollingK, thermalRollingSurfaceK, softnessIndex, radiusRaiseK;
    float grainThreshold, blisterThreshold, grainGamma, blisterGamma, grainGain, blisterGain, optimumTemp;
    /* TyreModelData */
    int32_t version;
    float Fz0, relaxationLength, rr0, rr1, rr_slip;
    float pressureSpringGain, pressureRRGain, pressureGainD, idealPressure, pressureRef;
    float Dx0, Dx1, lsMultX, lsExpX;   /* Tyre::getDX */
    PdCurve wearCurve;
    /* SCTM (Car/TyreModel.h:18-37) */
    float lsMultY, lsExpY, sctmLsMultX, sctmLsExpX, sctmFz0, maxSlip0, maxSlip1, asy, falloffSpeed;
    float speedSensitivity, camberGain, dcamber0, dcamber1, cfXmult, pressureCfGain, brakeDXMod;
    float dCamberBlend, combinedFactor;
    /* thermal */
    float surfaceTransfer, patchTransfer, patchCoreTransfer, internalCoreTransfer, coolFactorGain;
    float camberSpreadK;
    PdCurve performanceCurve;
    /* Tyre */
    float flatSpotK, explosionTemperature, pressureTemperatureGain, pressureStaticDefault;
    int32_t driven, useLoadForVKM;
} PdTyre;

typedef struct PdWing { /* Car/Wing.h:9-24 */
    float position[3], area, cdGain, clGain, angle, angleMult, yawGain;
    int32_t isVertical;
    PdCurve lutAOA_CL, lutAOA_CD;
} PdWing;

typedef struct PdEngine { /* Car/Engine.h:9-115 */
    PdCurve powerCurve, throttleResponseCurve;
    int32_t minimum, limiter, limiterCycles;
    float coast1, coast2, inertia, limiterMultiplier;
    float rpmDamageThreshold, rpmDamageK, turboBoostDamageThreshold, turboBoostDamageK, bovThreshold;
    float gasCoastOffset; int32_t coastEntryRpm;
    float overlapFreq, overlapGain, overlapIdealRPM;
    int32_t isEngineStallEnabled;
    float maxPowerRPM, maxTorqueRPM;
    /* engine.ini [THROTTLE_RESPONSE]: a second throttle map blended in with rpm / RPM_REFERENCE (Engine::getThrottleResponseGas, Engine.cpp:344-366) */
    PdCurve throttleResponseCurveMax; float throttleResponseCurveMaxRef; int32_t pad0;
} PdEngine;

typedef struct PdDrivetrain { /* Car/Drivetrain.h:75-96 */
    double gears[PD_MAX_GEARS]; int32_t nGears;
    int32_t tractionType, diffType;
    double finalRatio, diffPowerRamp, diffCoastRamp, diffPreLoad;
    double gearUpTime, gearDnTime, autoCutOffTime, controlsWindowGain, orgRpmWindow, damageRpmWindow;
    double clutchMaxTorque, clutchInertia, driveInertia, shaftInertiaL, shaftInertiaR;
    int32_t isShifterSupported, pad0;
} PdDrivetrain;

typedef struct PdAssists { /* AutoClutch / AutoBlip / AutoShifter */
    PdCurve upshiftProfile, downshiftProfile, blipProfile;
    float acRpmMin, acRpmMax, acClutchSpeed;
    int32_t acUseAutoOnStart, acUseAutoOnChange, acIsForced;
    double blipPerformTime; int32_t blipIsActive, blipIsElectronic;
    int32_t asChangeUpRpm, asChangeDnRpm; float asSlipThreshold, asGasCutoffTime; int32_t asIsActive, pad0;
} PdAssists;

typedef struct PdBrakeDisc { /* Car/BrakeSystem.h:15-25 */
    PdCurve perfCurve;
    float torqueK, coolTransfer, coolSpeedFactor;
} PdBrakeDisc;
typedef struct PdBrakes { /* Car/BrakeSystem.h:43-66 */
    float brakePower, brakePowerMultiplier, handBrakeTorque, frontBias, biasMin, biasMax;
    int32_t ebbInternal;          /* brakes.ini [EBB]: EBBMode::Internal -- front share follows the load distribution (BrakeSystem.cpp:95-118) */
    float ebbFrontMultiplier;
    int32_t hasTemps;             /* brakes.ini [TEMPS_FRONT] + [TEMPS_REAR]: disc temperatures scale the brake torque (BrakeSystem.cpp:151-168) */
    PdBrakeDisc disc[PD_NUM_WHEELS];
} PdBrakes;

/* indices into PdCarParams::scoring (Car/ScoringSystem.cpp:50-73, same order) */
enum {
    PD_SV_SmoothSteerSpeed = 0, PD_SV_MinBonusSpeed, PD_SV_MaxBonusSpeed, PD_SV_StallRpm, PD_SV_DirectionThreshold,
    PD_SV_OutOfTrackThreshold, PD_SV_ApproachDistance, PD_SV_CriticalDistance, PD_SV_TravelBonus,
    PD_SV_TravelSplineBonus, PD_SV_DriftBonus, PD_SV_SpeedBonus, PD_SV_ThrottleBonus, PD_SV_EngineRpmBonus,
    PD_SV_DirectionBonus, PD_SV_DirectionPenalty, PD_SV_ObstApproachPenalty, PD_SV_CollisionPenalty,
    PD_SV_OffTrackPenalty, PD_SV_GearGrindPenalty, PD_SV_StallPenalty
};

#define PD_MAX_COLLIDER_VERTS 128     /* bundled cars: 50 .. 113 vertices, 96 .. 178 triangles (indices fit a byte) */
#define PD_MAX_COLLIDER_TRIS 192

typedef struct PdCarParams {
    /* Car (car.ini) */
    float mass, chassisMass, chassisInertia[3], tankMass, tankInertia[3];
    float fuelTankPos[3];
    float steerLock, steerRatio, steerLinearRatio;
    float fuelKG; double fuelConsumptionK, maxFuel; float requestedFuel;
    int32_t framesToSleep;
    float waterTmass, waterCoolSpeedK, waterCoolFactor, waterHeatFactor;
    float baseCarHeight;      /* Car::getBaseCarHeight (Car.cpp:1360-1365) */
    /* fixed joint tank(b0) <-> chassis(b1) (Car.cpp:51) */
    float tankOffset[3], tankQrel[4];
    PdStrut strut[2];
    PdAxle axle;
    float arbK[2];            /* AntirollBar::k front, rear */
    PdBrakes brakes;
    PdTyre tyre[PD_NUM_WHEELS];
    int32_t nWings; PdWing wing[PD_MAX_WINGS];
    PdEngine engine;
    PdDrivetrain drivetrain;
    PdAssists assists;
    /* probes / look-ahead (cfg/sim.ini, Car.cpp:288-316) */
    int32_t nProbes; float probeDir[PD_MAX_PROBES][3]; float probeLength[PD_MAX_PROBES];
    int32_t lookAheadCount; float lookAheadStep;
    /* scoring (process-global singleton in the reference) */
    float scoring[PD_NUM_SCORING_VARS];
    int32_t teleportOnCollision, teleportOnBadLocation, teleportMode;
    /* Simulator (Sim/Simulator.cpp:34-42,346-349) + world (PhysicsEngineODE.cpp:23-28) */
    float ambientTemperature, roadTemperature, airDensity;
    float fuelConsumptionRate, tyreConsumptionRate, mechanicalDamageRate;
    int32_t allowTyreBlankets;
    float gravityY, worldERP, worldCFM;
    /* colliders of the chassis: the box of colliders.ini [COLLIDER_0] (CarColliderManager.cpp:12-34,
     * category CAR, collides with TRACK meshes) and the hull mesh of collider.bin (Car.cpp:318-379, collides with WALL
     * meshes).  Mesh vertices are stored chassis-local, the geom offset (Car::getGraphicsOffsetMatrix at construction:
     * GRAPHICS_OFFSET, identity rotation) already added. */
    int32_t hasBoxCollider; float boxCentre[3]; float boxSize[3];
    int32_t nColliderVerts, nColliderTris;
    float colliderMin[3], colliderMax[3];                  /* chassis-local bounds of the hull mesh */
    float colliderVerts[PD_MAX_COLLIDER_VERTS][3];
    uint8_t colliderTris[PD_MAX_COLLIDER_TRIS][4];         /* vertex indices i0, i1, i2, 0 */
    float colliderTriBounds[PD_MAX_COLLIDER_TRIS][6];      /* chassis-local box of each hull triangle, grown by 1e-4: min xyz, max xyz (a filter only) */
    float colliderTriSphere[PD_MAX_COLLIDER_TRIS][4];      /* bounding sphere of each hull triangle: centre xyz, radius grown by 1e-4 (a filter only) */
    /* the other bundled cars: double-wishbone axles and turbochargers.  dw[w] is filled for the wheels of a DWB
     * axle (strut[] / axle stay zero for that axle), topology = PD_TOPO_* */
    int32_t topology, nTurbos;
    PdDW dw[PD_NUM_WHEELS];
    PdTurbo turbo[PD_MAX_TURBOS];
    PdHeave heave[2];          /* front, rear (only between two double-wishbone corners, Car.cpp:131-146) */
} PdCarParams;

/* ---- track (Sim/Track.cpp) ---- */
typedef struct PdSurface { /* Sim/Surface.h:7-24 */
    float gripMod, damping, sinHeight, sinLength, granularity, dirtAdditiveK;
    uint32_t collisionCategory, sectorID;
    uint32_t isValidTrack, isPitlane, pad0, pad1;
} PdSurface;

typedef struct PdFatPoint { float best[3], left[3], right[3], center[3], forwardDir[3]; } PdFatPoint; /* Track.h:19-25 */

typedef struct PdTrackInfo {
    int32_t nSurfaces, nTris, nNodes, nFatPoints, nSplineNodes, interpolateStep, closedLoop, pad0;
    float computedTrackLength, computedTrackWidth, dynamicGripLevel, hashCellSize;
} PdTrackInfo;

/* uniform x-z grid over the track's boundary polylines and spline points: an INDEX only (it prunes the
 * candidates of Track::rayCastTrackBounds / getPointIdAtLocation, it does not change their results) */
typedef struct PdBoundGrid {
    float ox, oz, cell, invCell;
    int32_t nx, nz, pad0, pad1;
} PdBoundGrid;

#ifdef __cplusplus
}
#endif
#endif /* PD_PARAMS_H */
