/*
 * pd_state.h -- canonical per-car state record of the batched Car::step path.
 *
 * One record = everything the reference keeps between two ticks for ONE car of the
 * demo-car topology: 7 rigid bodies, 4 tyres with
 * their 12x3 thermal patch grid, drivetrain / engine / assist state machines, track
 * locator and scoring state.  On the host (get / set state, parity tests, the oracle harness) a record is
 * the flat array of PD_STATE_WORDS 32-bit words laid out by the X-macros below (doubles = 2 words at an even
 * word offset).  On the device the library keeps either an array of such records (stride PD_STATE_STRIDE,
 * staged through shared memory by the 4-lanes-per-car kernel) or a tiled structure-of-arrays (thread-per-car
 * kernel); see projectd_core_b200/csrc/pd_state_io.h.
 *
 * Reference anchors for every group are given next to the list that defines it
 * (paths relative to the reference tree, src/ProjectD/...).
 *
 * Field kinds:  F = float (1 word)   I = int32 (1 word)   D = double (2 words: lo, hi)
 */
#ifndef PD_STATE_H
#define PD_STATE_H

#include <stdint.h>

#define PD_NUM_BODIES   8
#define PD_BODY_CHASSIS 0  /* Car::body              Car/Car.cpp:38        */
#define PD_BODY_TANK    1  /* Car::fuelTankBody      Car/Car.cpp:39,49-51  */
#define PD_BODY_HUB0    2  /* SuspensionStrut::hub   (LF)  SuspensionStrut.cpp:132 */
#define PD_BODY_STRUT0  3  /* SuspensionStrut::strutBody (LF) SuspensionStrut.cpp:135 */
#define PD_BODY_HUB1    4  /* (RF) */
#define PD_BODY_STRUT1  5  /* (RF) */
#define PD_BODY_AXLE    6  /* Car::rigidAxle         Car/Car.cpp:66, SuspensionAxle.cpp:46 */
/* cars with double-wishbone suspensions (SuspensionDW::hub, SuspensionDW.cpp:143): a DWB front axle keeps its hubs in the
 * HUB0 / HUB1 slots (the strut-body slots stay zero), a DWB rear axle keeps its two hubs in slot 6 (instead of the rigid
 * axle) and slot 7.  Slots a topology does not have are zero in every record and are never touched by its kernels. */
#define PD_BODY_HUB2    6  /* (LR, DWB rear) */
#define PD_BODY_HUB3    7  /* (RR, DWB rear) */

#define PD_NUM_WHEELS   4   /* LF, RF, LR, RR (Car/Car.cpp:72) */
#define PD_THERMAL_STRIPES  3   /* Tyre.cpp:41 thermalModel->init(car, 12, 3) */
#define PD_THERMAL_ELEMENTS 12
#define PD_THERMAL_PATCHES  36
#define PD_MAX_PROBES   10  /* CarState::MaxProbes    Car/CarState.h:48 */
#define PD_LOOKAHEAD    5   /* CarState::MaxLookAhead Car/CarState.h:51 */
#define PD_OBS_DIM      24  /* pyprojectd/projectd_env.py:237-275 */

/* ---- rigid body: dxBody pos/q/R/lvel/avel (ODE 0.16.3 objects; RigidBodyODE.cpp:131-231).
 *      R is kept beside q exactly as ODE keeps both: after dBodySetRotation (teleports) R is the
 *      orthogonalised input while q comes from the raw input, and they only re-synchronise at the next
 *      dxStepBody.  (axx,axy,axz) = body x axis in world = mat44f row 1 = first COLUMN of ODE's R. ---- */
#define PD_BODY_FIELDS(X) \
    X(F, px) X(F, py) X(F, pz) \
    X(F, qw) X(F, qx) X(F, qy) X(F, qz) \
    X(F, vx) X(F, vy) X(F, vz) \
    X(F, wx) X(F, wy) X(F, wz) \
    X(F, axx) X(F, axy) X(F, axz) X(F, ayx) X(F, ayy) X(F, ayz) X(F, azx) X(F, azy) X(F, azz)

/* ---- tyre: TyreStatus (Car/TyreStatus.h:5-45) + Tyre runtime members (Car/Tyre.h:88-106) ---- */
#define PD_TYRE_FIELDS(X) \
    /* doubles first: the group starts on an even word, so they are 8-byte aligned in a record */ \
    X(D, virtualKM) X(D, flatSpot) X(D, phase) \
    X(F, depth) X(F, load) X(F, camberRAD) X(F, slipAngleRAD) X(F, slipRatio) \
    X(F, angularVelocity) X(F, Fy) X(F, Fx) X(F, Mz) X(I, isLocked) \
    X(F, slipFactor) X(F, ndSlip) X(F, distToGround) X(F, Dy) X(F, Dx) X(F, D) \
    X(F, dirtyLevel) X(F, rollingResistence) X(F, thermalInput) X(F, feedbackTorque) \
    X(F, loadedRadius) X(F, effectiveRadius) X(F, liveRadius) \
    X(F, pressureStatic) X(F, pressureDynamic) X(F, inflation) \
    X(F, wearMult) \
    X(F, oldAngularVelocity) X(F, localMX) \
    X(F, contactX) X(F, contactY) X(F, contactZ) \
    X(F, normalX) X(F, normalY) X(F, normalZ) \
    X(I, surfaceId) X(I, hasContact) \
    X(F, totalHubVelocity) X(F, slidingVelocityX) X(F, slidingVelocityY) \
    X(F, brakeTorque) X(F, handBrakeTorque) \
    X(F, suspTravel) X(F, suspDamperSpeed) \
    /* TyreThermalModel (Car/TyreThermalModel.h:36-49) */ \
    X(F, coreTemp) X(F, practicalTemp) X(F, thermalMultD) \
    X(I, tyrePad)   /* keeps the tyre group an even number of words */

/* ---- car-level (Car/Car.h:187-246, AutoClutch.h, AutoBlip.h, AutoShifter.h, GearChanger.h,
 *      Drivetrain.h:98-145, Engine.h:96-115, ScoringSystem.h:47-66, Track.h:76,79) ---- */
#define PD_CAR_FIELDS(X) \
    /* doubles first (8-byte aligned inside the record): fuel, AutoBlip, Drivetrain (Drivetrain.h:98-145), Engine */ \
    X(D, fuel) X(D, blipStartTime) X(D, reqTimeAcc) X(D, reqTimeout) \
    X(D, engineVel) X(D, driveVel) X(D, shaftLVel) X(D, shaftRVel) X(D, rootVel) \
    X(D, locClutch) X(D, lastRatio) X(D, cutOff) \
    X(D, validShiftRPMWindow) X(D, currentClutchTorque) X(D, outTorque) \
    /* CarControls (Car/CarControls.h:9-20) as last written by the caller / the assists */ \
    X(F, ctlSteer) X(F, ctlClutch) X(F, ctlBrake) X(F, ctlHandBrake) X(F, ctlGas) \
    X(I, ctlRequestedGear) X(I, ctlGearUp) X(I, ctlGearDn) X(I, smoothSteer) \
    X(F, smoothSteerValue) X(F, finalSteerAngleSignal) \
    X(F, lastVelX) X(F, lastVelY) X(F, lastVelZ) \
    X(F, accGX) X(F, accGY) X(F, accGZ) \
    X(I, sleepingFrames) X(F, waterT) X(F, speed) \
    X(I, collisionFlag) X(I, outOfTrackFlag) \
    /* track locator */ \
    X(I, nearestTrackPointId) X(I, oldTrackPointId) X(I, splinePointId) \
    X(F, lastTrackPointTimestamp) X(F, trackLocation) X(F, oldTrackLocation) \
    X(F, bodyVsTrack) X(F, velocityVsTrack) \
    X(F, pointCacheX) X(F, pointCacheY) X(F, pointCacheZ) \
    /* AutoClutch */ \
    X(F, acSeqTime) X(I, acSeqDone) X(I, acSeqProfile) X(F, acClutchValueSignal) \
    /* AutoBlip / AutoShifter / GearChanger */ \
    X(F, gasCutoff) X(I, lastGearUp) X(I, lastGearDn) \
    /* Drivetrain */ \
    X(I, reqRequest) X(I, reqGear) \
    X(I, currentGear) X(I, isGearGrinding) X(I, clutchOpenState) \
    /* Engine */ \
    X(I, limiterOn) X(F, lifeLeft) X(F, fuelPressure) X(F, gasUsage) \
    /* ScoringSystem */ \
    X(I, drifting) X(I, driftExtreme) X(I, driftInvalid) \
    X(F, currentDriftAngle) X(F, currentSpeedMultiplier) X(F, lastDriftDirection) \
    X(F, driftStraightTimer) X(F, instantDriftDelta) X(F, instantDrift) X(F, driftPoints) \
    X(I, driftComboCounter) X(F, stepReward) X(F, totalReward) X(F, prevEpisodeReward) \
    X(I, oldPointId) X(I, oldSplinePointId) \
    /* batched-env bookkeeping (no reference counterpart: episode statistics, NaN guard) */ \
    X(I, episodeSteps) X(I, nanFlag) \
    /* TyreThermalPatch::inputT starts at ambient (TyreThermalModel.cpp:40) and is zero after the first step */ \
    X(I, thermalPrimed) \
    /* PhysicsEngineODE::currentFrame (PhysicsEngineODE.cpp:228-244): collisions against the static meshes are tested on odd frames */ \
    X(I, physFrame) \
    /* Car::damageZoneLevel[5] (Car.h:204; front, rear, left, right, max): raised by wall contacts in Car::onCollisionCallback \
     * (Car.cpp:980-999), read by ScoringSystem::validateDrift (a tick with new damage invalidates the drift); + 1 pad word */ \
    X(F, damageZone0) X(F, damageZone1) X(F, damageZone2) X(F, damageZone3) X(F, damageZone4) X(I, carPad) \
    /* Turbo::rotation of up to three turbochargers (Turbo.h, Turbo.cpp:11-40) and Engine::status.turboBoost (Engine.cpp:368-384) */ \
    X(F, turboRot0) X(F, turboRot1) X(F, turboRot2) X(F, turboBoost) \
    /* BrakeDisc::t of the four discs (BrakeSystem.h:15-25, BrakeSystem.cpp:151-168): only cars with [TEMPS_FRONT] / [TEMPS_REAR] move them */ \
    X(F, brakeDiscT0) X(F, brakeDiscT1) X(F, brakeDiscT2) X(F, brakeDiscT3)

/* ---------------------------------------------------------------------------------------- */
#define PD__W_F 1
#define PD__W_I 1
#define PD__W_D 2
#define PD__COUNT(kind, name) + PD__W_##kind

#define PD_BODY_WORDS   (0 PD_BODY_FIELDS(PD__COUNT))
#define PD_TYRE_SCALAR_WORDS (0 PD_TYRE_FIELDS(PD__COUNT))
#define PD_TYRE_WORDS   (PD_TYRE_SCALAR_WORDS + PD_THERMAL_PATCHES)
#define PD_CAR_SCALAR_WORDS  (0 PD_CAR_FIELDS(PD__COUNT))
#define PD_CAR_WORDS    (PD_CAR_SCALAR_WORDS + PD_MAX_PROBES + PD_LOOKAHEAD)

/* word offsets of the groups inside one record */
#define PD_OFF_BODY(b)   ((b) * PD_BODY_WORDS)
#define PD_OFF_TYRE(w)   (PD_NUM_BODIES * PD_BODY_WORDS + (w) * PD_TYRE_WORDS)
#define PD_OFF_TYRE_PATCH(w) (PD_OFF_TYRE(w) + PD_TYRE_SCALAR_WORDS)
#define PD_OFF_CAR       (PD_OFF_TYRE(PD_NUM_WHEELS))
#define PD_OFF_PROBES    (PD_OFF_CAR + PD_CAR_SCALAR_WORDS)
#define PD_OFF_LOOKAHEAD (PD_OFF_PROBES + PD_MAX_PROBES)
#define PD_STATE_WORDS   (PD_OFF_CAR + PD_CAR_WORDS)
/* record stride of the array-of-records device layout: a multiple of 4 words (16-byte bulk copies) whose value
 * mod 32 (= 28: its multiples run through 0, 28, 24, 20, 16, 12, 8, 4) spreads the same word of 8 consecutive records over 8 different shared-memory bank groups */
#define PD_STATE_STRIDE  668

/* per-field word offsets inside their group: PD_BODY_o_px, PD_TYRE_o_load, PD_CAR_o_fuel ... */
#define PD__ENUM_B(kind, name) PD_BODY_o_##name, PD_BODY_e_##name = PD_BODY_o_##name + PD__W_##kind - 1,
#define PD__ENUM_T(kind, name) PD_TYRE_o_##name, PD_TYRE_e_##name = PD_TYRE_o_##name + PD__W_##kind - 1,
#define PD__ENUM_C(kind, name) PD_CAR_o_##name,  PD_CAR_e_##name  = PD_CAR_o_##name  + PD__W_##kind - 1,
enum { PD_BODY_FIELDS(PD__ENUM_B) PD_BODY__end };
enum { PD_TYRE_FIELDS(PD__ENUM_T) PD_TYRE__end };
enum { PD_CAR_FIELDS(PD__ENUM_C)  PD_CAR__end };

#endif /* PD_STATE_H */
