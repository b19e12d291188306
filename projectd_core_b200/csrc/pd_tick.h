/*
 * pd_tick.h -- one full simulator tick for one car: Simulator::step (Sim/Simulator.cpp:168-201) =
 * stepPreCacheValues + Car::step (Car/Car.cpp:421-553) + stepComponents (:638-681) + dWorldStep +
 * Car::postStep (:685-713), and the teleport / reset path (Car.cpp:1240-1358).
 */
#pragma once
#include "pd_solver2.h"
#include "pd_contacts.h"
#include "../../include/pd_batch.h"

namespace pd {

/* SuspensionBase::baseCFM of the corners that follow Car::step's ERP / CFM switch (Car.cpp:426-451): 1e-7 in every suspension class that sets it;
 * a multilink corner ignores the switch (its setERPCFM is empty), so it is not asked */
template <int TOPO> PD_HD float susp_base_cfm(const PdCarParams& P) { return PD_TOPO_FRONT_DW(TOPO) ? P.dw[0].baseCFM : P.strut[0].baseCFM; }
/* topo: the car's suspension topology (PD_TOPO_*; a compile-time constant inside the tick kernels) */
PD_HD void set_body_mass(Body* b, const PdCarParams& P, int topo) {
    b[PD_BODY_CHASSIS].mass = P.chassisMass; b[PD_BODY_CHASSIS].I = v3(P.chassisInertia[0], P.chassisInertia[1], P.chassisInertia[2]);
    b[PD_BODY_TANK].mass = P.tankMass; b[PD_BODY_TANK].I = v3(P.tankInertia[0], P.tankInertia[1], P.tankInertia[2]);
    for (int s = 0; s < 2; ++s) {
        if (PD_TOPO_FRONT_DW(topo)) { b[PD_BODY_HUB0 + 2 * s].mass = P.dw[s].hubMass; b[PD_BODY_HUB0 + 2 * s].I = v3(P.dw[s].hubInertia[0], P.dw[s].hubInertia[1], P.dw[s].hubInertia[2]); continue; }
        b[PD_BODY_HUB0 + 2 * s].mass = P.strut[s].hubMass; b[PD_BODY_HUB0 + 2 * s].I = v3(P.strut[s].hubInertia[0], P.strut[s].hubInertia[1], P.strut[s].hubInertia[2]);
        b[PD_BODY_STRUT0 + 2 * s].mass = P.strut[s].strutMass; b[PD_BODY_STRUT0 + 2 * s].I = v3(P.strut[s].strutInertia[0], P.strut[s].strutInertia[1], P.strut[s].strutInertia[2]);
    }
    if (PD_TOPO_REAR_DW(topo)) {
        for (int s = 0; s < 2; ++s) { b[PD_BODY_HUB2 + s].mass = P.dw[2 + s].hubMass; b[PD_BODY_HUB2 + s].I = v3(P.dw[2 + s].hubInertia[0], P.dw[2 + s].hubInertia[1], P.dw[2 + s].hubInertia[2]); }
    } else { b[PD_BODY_AXLE].mass = P.axle.axleMass; b[PD_BODY_AXLE].I = v3(P.axle.axleInertia[0], P.axle.axleInertia[1], P.axle.axleInertia[2]); }
}
/* SuspensionDW::attach (SuspensionDW.cpp:157-161): hub rotation = the chassis' world matrix, hub at the reference point */
PD_HD void dw_attach(const PdDW& D, const Body& C, Body& H) {
    set_rotation(C.fr.ax, C.fr.ay, C.fr.az, H.fr.ax, H.fr.ay, H.fr.az, H.q);
    H.fr.p = to_world(C.fr, v3(D.refPoint[0], D.refPoint[1], D.refPoint[2]));
}

/* Car::getBetaRad (Car.cpp:1472-1484) */
PD_HD float beta_rad(const Body& C) {
    V3 vel = irot(C.fr, C.v);
    const float fLen = len(vel);
    if (fLen != 0.0f) vel.x /= fLen;
    if (vel.x <= -1.0f || vel.x >= 1.0f) return 1.5707964f;
    return m_asin(vel.x);
}

/* state of a freshly constructed car (Car::init, Car.cpp:31-223, and the constructors it runs): chassis at the
 * origin with identity rotation, suspension bodies attached, everything else at its default */
template <class SVX> PD_HDN void car_init_state(const PdCarParams& P, const SVX& sv) {
    for (int w = 0; w < PD_STATE_WORDS; ++w) sv.i(w, 0);
    const int topo = P.topology;
    Body bod[PD_NUM_BODIES]; set_body_mass(bod, P, topo);
    for (int i = 0; i < PD_NUM_BODIES; ++i) {
        Body& b = bod[i]; b.fr.p = v3(0, 0, 0); b.fr.ax = v3(1, 0, 0); b.fr.ay = v3(0, 1, 0); b.fr.az = v3(0, 0, 1);
        b.q.w = 1; b.q.x = b.q.y = b.q.z = 0; b.v = v3(0, 0, 0); b.w = v3(0, 0, 0);
    }
    Body& C = bod[PD_BODY_CHASSIS];
    bod[PD_BODY_TANK].fr.p = v3(P.fuelTankPos[0], P.fuelTankPos[1], P.fuelTankPos[2]);
    for (int s = 0; s < 2; ++s) {
        if (PD_TOPO_FRONT_DW(topo)) { bod[PD_BODY_HUB0 + 2 * s].fr.p = to_world(C.fr, v3(P.dw[s].refPoint[0], P.dw[s].refPoint[1], P.dw[s].refPoint[2])); continue; }
        const PdStrut& S = P.strut[s]; Body& H = bod[PD_BODY_HUB0 + 2 * s]; Body& B = bod[PD_BODY_STRUT0 + 2 * s];
        H.fr.p = to_world(C.fr, v3(S.refPoint[0], S.refPoint[1], S.refPoint[2]));
        const V3 vCarStrut = to_world(C.fr, v3(S.carStrut[0], S.carStrut[1], S.carStrut[2]));
        const V3 vTyreStrut = to_world(H.fr, v3(S.tyreStrut[0], S.tyreStrut[1], S.tyreStrut[2]));
        const V3 vNorm = norm(vTyreStrut - vCarStrut);
        const V3 vM3 = C.fr.az * -1.0f;
        const V3 vM3N = cross(vM3, vNorm);
        const V3 vM3NN = norm(cross(vM3N, vNorm));
        set_rotation(vM3NN, vM3N * -1.0f, vNorm * -1.0f, B.fr.ax, B.fr.ay, B.fr.az, B.q);
        B.fr.p = (vNorm * S.strutBodyLength) * 0.5f + vCarStrut;
    }
    if (PD_TOPO_REAR_DW(topo)) { for (int s = 0; s < 2; ++s) bod[PD_BODY_HUB2 + s].fr.p = to_world(C.fr, v3(P.dw[2 + s].refPoint[0], P.dw[2 + s].refPoint[1], P.dw[2 + s].refPoint[2])); }
    else bod[PD_BODY_AXLE].fr.p = to_world(C.fr, v3(P.axle.axleBasePos[0], P.axle.axleBasePos[1], P.axle.axleBasePos[2]));
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body_rt(topo, i)) store_body(sv, i, bod[i]);
    for (int w = 0; w < PD_NUM_WHEELS; ++w) { /* Tyre::Tyre + setCompound(0) + reset (Tyre.cpp:15-26,343-425) */
        const int o = PD_OFF_TYRE(w);
        sv.f(o + PD_TYRE_o_pressureStatic, P.tyre[w].pressureStaticDefault); sv.f(o + PD_TYRE_o_pressureDynamic, P.tyre[w].pressureRef);
        sv.i(o + PD_TYRE_o_isLocked, 1); sv.f(o + PD_TYRE_o_inflation, 1);
        sv.i(o + PD_TYRE_o_surfaceId, -1);
        sv.f(o + PD_TYRE_o_coreTemp, P.ambientTemperature); sv.f(o + PD_TYRE_o_thermalMultD, 1.0f);
        for (int p = 0; p < PD_THERMAL_PATCHES; ++p) sv.f(PD_OFF_TYRE_PATCH(w) + p, P.ambientTemperature);
    }
    const int o = PD_OFF_CAR;
    sv.i(o + PD_CAR_o_ctlRequestedGear, -1);
    sv.d(o + PD_CAR_o_fuel, P.requestedFuel);
    sv.f(o + PD_CAR_o_pointCacheY, -10000.0f);
    sv.i(o + PD_CAR_o_acSeqDone, 1);
    sv.d(o + PD_CAR_o_reqTimeout, 200); sv.i(o + PD_CAR_o_reqGear, -1);
    sv.d(o + PD_CAR_o_locClutch, 1.0); sv.d(o + PD_CAR_o_lastRatio, -1.0);
    sv.i(o + PD_CAR_o_currentGear, 1);
    sv.d(o + PD_CAR_o_validShiftRPMWindow, P.drivetrain.orgRpmWindow);
    sv.f(o + PD_CAR_o_lifeLeft, 1000.0f); sv.f(o + PD_CAR_o_fuelPressure, 1.0f);
}

/* where Car::teleportToSpline puts the chassis: forceRotation(heading of the spline point) + forcePosition(its centre, dropped
 * onto the ground by one ray from 10 m above) (Car.cpp:1325-1340,1275-1308,1240-1273) */
PD_HD void teleport_chassis_pose(const PdCarParams& P, const TrackDev& T, int pointId, V3& oax, V3& oay, V3& oaz, Quat& q, V3& pos) {
    const PdFatPoint& pt = T.fat[pointId];
    {
        const V3 heading = v3(pt.forwardDir[0], pt.forwardDir[1], pt.forwardDir[2]);
        const V3 ihed = heading * -1.0f;
        const float vM13 = ihed.x, vM11 = -ihed.z, vM12 = 0;
        const float v6 = sqrtf((vM12 * vM12) + (vM11 * vM11) + (vM13 * vM13));
        const float s = 1.0f / v6;
        const V3 ax = v3(vM11 * s, vM12 * s, vM13 * s), ay = v3(0, 1, 0), az = v3(-ihed.x, -ihed.y, -ihed.z);
        set_rotation(ax, ay, az, oax, oay, oaz, q);
    }
    V3 bodyPos = v3(pt.center[0], pt.center[1], pt.center[2]);
    {
        const RayHit hit = ray_cast_down(T, bodyPos + v3(0, 10, 0), 1000.0f);
        if (hit.hit) bodyPos.y = hit.pos.y;
        bodyPos.y += (P.baseCarHeight + 0.0f + 0.01f);
    }
    pos = bodyPos;
}

/* teleport: Car::teleportToSpline -> forceRotation + forcePosition (Car.cpp:1325-1340,1275-1308,1240-1273) */
template <class SVX> PD_HDN void car_teleport_to_point(const PdCarParams& P, const TrackDev& T, const SVX& sv, int pointId, double physicsTime) {
    const int topo = P.topology;
    Body bod[PD_NUM_BODIES]; set_body_mass(bod, P, topo);
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body_rt(topo, i)) load_body(sv, i, bod[i]);
    CarS c; load_car(sv, c);
    Body& C = bod[PD_BODY_CHASSIS]; Body& Tk = bod[PD_BODY_TANK];
    V3 bodyPos;
    teleport_chassis_pose(P, T, pointId, C.fr.ax, C.fr.ay, C.fr.az, C.q, bodyPos);
    Tk.fr.ax = C.fr.ax; Tk.fr.ay = C.fr.ay; Tk.fr.az = C.fr.az; Tk.q = C.q;
    /* Car::reset (Car.cpp:385-410) */
    c.waterT = 60; c.fuel = P.requestedFuel;
    c.damageZone0 = 0; c.damageZone1 = 0; c.damageZone2 = 0; c.damageZone3 = 0; c.damageZone4 = 0;      /* Car.cpp:403-407 */
    c.collisionFlag = 0; c.outOfTrackFlag = 0;
    c.lastTrackPointTimestamp = (float)physicsTime;
    c.nearestTrackPointId = 0; c.oldTrackPointId = 0; c.splinePointId = 0; c.trackLocation = 0; c.oldTrackLocation = 0;
    c.prevEpisodeReward = c.totalReward; c.totalReward = 0; c.stepReward = 0; c.oldPointId = 0; c.oldSplinePointId = 0;
    c.episodeSteps = 0;
    body_stop(C); C.fr.p = bodyPos;
    Tk.fr.p = to_world(C.fr, v3(P.fuelTankPos[0], P.fuelTankPos[1], P.fuelTankPos[2]));
    /* susp->stop(); susp->attach() */
    for (int s = 0; s < 2; ++s) { /* SuspensionStrut::setPositions (SuspensionStrut.cpp:188-223) */
        if (PD_TOPO_FRONT_DW(topo)) { Body& H = bod[PD_BODY_HUB0 + 2 * s]; body_stop(H); dw_attach(P.dw[s], C, H); continue; }
        const PdStrut& S = P.strut[s]; Body& H = bod[PD_BODY_HUB0 + 2 * s]; Body& B = bod[PD_BODY_STRUT0 + 2 * s];
        body_stop(H);
        set_rotation(C.fr.ax, C.fr.ay, C.fr.az, H.fr.ax, H.fr.ay, H.fr.az, H.q);   /* hub->setRotation(mxBody) */
        H.fr.p = to_world(C.fr, v3(S.refPoint[0], S.refPoint[1], S.refPoint[2]));
        const V3 vCarStrut = to_world(C.fr, v3(S.carStrut[0], S.carStrut[1], S.carStrut[2]));
        const V3 vTyreStrut = to_world(H.fr, v3(S.tyreStrut[0], S.tyreStrut[1], S.tyreStrut[2]));
        const V3 vNorm = norm(vTyreStrut - vCarStrut);
        const V3 vM3 = C.fr.az * -1.0f;
        const V3 vM3N = cross(vM3, vNorm);
        const V3 vM3NN = norm(cross(vM3N, vNorm));
        set_rotation(vM3NN, vM3N * -1.0f, vNorm * -1.0f, B.fr.ax, B.fr.ay, B.fr.az, B.q);
        B.fr.p = (vNorm * S.strutBodyLength) * 0.5f + vCarStrut;
        /* NB SuspensionStrut::stop() stops only the hub; the strut body keeps its velocity (reference behaviour) */
    }
    if (PD_TOPO_REAR_DW(topo)) { for (int s = 0; s < 2; ++s) { Body& H = bod[PD_BODY_HUB2 + s]; body_stop(H); dw_attach(P.dw[2 + s], C, H); } }
    else {
        Body& A = bod[PD_BODY_AXLE]; body_stop(A);
        set_rotation(C.fr.ax, C.fr.ay, C.fr.az, A.fr.ax, A.fr.ay, A.fr.az, A.q);
        A.fr.p = to_world(C.fr, v3(P.axle.axleBasePos[0], P.axle.axleBasePos[1], P.axle.axleBasePos[2]));
    }
    /* drivetrain->reset() (Drivetrain.cpp:154-167) */
    c.clutchOpenState = 1; c.rootVel = 0; c.engineVel = 0; c.shaftLVel = 0; c.shaftRVel = 0; c.driveVel = 0;
    c.reqRequest = 0; c.validShiftRPMWindow = P.drivetrain.orgRpmWindow; c.lifeLeft = 1000.0f;
    c.brakeDiscT0 = P.ambientTemperature; c.brakeDiscT1 = P.ambientTemperature; c.brakeDiscT2 = P.ambientTemperature; c.brakeDiscT3 = P.ambientTemperature;   /* brakeSystem->reset() (BrakeSystem.cpp:73-80) */
    c.turboRot0 = 0; c.turboRot1 = 0; c.turboRot2 = 0;          /* Engine::reset -> Turbo::reset (Engine.cpp:160-166); status.turboBoost keeps its last value */
    /* tyres reset (Tyre.cpp:393-425) */
    for (int w = 0; w < PD_NUM_WHEELS; ++w) {
        const int o = PD_OFF_TYRE(w);
        sv.f(o + PD_TYRE_o_slipAngleRAD, 0); sv.f(o + PD_TYRE_o_slipRatio, 0); sv.f(o + PD_TYRE_o_angularVelocity, 0);
        sv.f(o + PD_TYRE_o_Fy, 0); sv.f(o + PD_TYRE_o_Fx, 0); sv.f(o + PD_TYRE_o_Mz, 0); sv.i(o + PD_TYRE_o_isLocked, 1);
        sv.f(o + PD_TYRE_o_inflation, 1); sv.d(o + PD_TYRE_o_flatSpot, 0);
        sv.f(o + PD_TYRE_o_dirtyLevel, 0); sv.f(o + PD_TYRE_o_feedbackTorque, 0); sv.d(o + PD_TYRE_o_virtualKM, 0);
        sv.f(o + PD_TYRE_o_coreTemp, P.ambientTemperature); sv.d(o + PD_TYRE_o_phase, 0);
        for (int p = 0; p < PD_THERMAL_PATCHES; ++p) sv.f(PD_OFF_TYRE_PATCH(w) + p, P.ambientTemperature);
    }
    /* drivetrain->setCurrentGear(1, true) */
    c.isGearGrinding = 0; c.currentGear = 1;
    body_stop(C); body_stop(Tk);
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body_rt(topo, i)) store_body(sv, i, bod[i]);
    store_car(sv, c);
}

/* Car::updateLookAhead (Car.cpp:775-798) */
PD_HD void post_lookahead(const PdCarParams& P, const TrackDev& T, const Body& C, CarS& c) {
    { /* updateLookAhead (Car.cpp:775-798) */
        const V3 up = v3(0, 1, 0);
        const V3 curTrackDir = track_direction_at_distance(T, c.trackLocation);
        const V3 bodyFrontDir = norm(C.fr.az);
        const float driveDir = signf_(dot(bodyFrontDir, curTrackDir));
        for (int i = 0; i < P.lookAheadCount && i < PD_LOOKAHEAD; ++i) {
            const float distanceNorm = c.trackLocation + ((P.lookAheadStep * (float)(i + 1)) / T.info.computedTrackLength) * driveDir;
            const V3 dir = track_direction_at_distance(T, distanceNorm);
            c.lookAhead[i] = m_atan2(dot(cross(dir, curTrackDir), up), dot(curTrackDir, dir));
        }
    }
}

/* the same, look-ahead points dealt to the four lanes of a quad (point i on lane i & 3), results exchanged */
template <class Ex> PD_HD void post_lookahead_quad(const PdCarParams& P, const TrackDev& T, const Body& C, CarS& c, Ex& ex) {
    const V3 up = v3(0, 1, 0);
    const V3 curTrackDir = track_direction_at_distance(T, c.trackLocation);
    const V3 bodyFrontDir = norm(C.fr.az);
    const float driveDir = signf_(dot(bodyFrontDir, curTrackDir));
    const int cnt = (P.lookAheadCount < PD_LOOKAHEAD) ? P.lookAheadCount : PD_LOOKAHEAD;
    float mine[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {
        const int i = ex.lane + 4 * k;
        if (i >= cnt) continue;
        if (ex.nhalf == 2) continue;                   /* with a helper quad the twins take one point each (below) */
        const float distanceNorm = c.trackLocation + ((P.lookAheadStep * (float)(i + 1)) / T.info.computedTrackLength) * driveDir;
        const V3 dir = track_direction_at_distance(T, distanceNorm);
        mine[k] = m_atan2(dot(cross(dir, curTrackDir), up), dot(curTrackDir, dir));
    }
    if (ex.nhalf == 2) {
        /* both twins run the same code on their own point (lane and lane + 4), then swap */
        const int i = ex.lane + 4 * ex.half;
        float v = 0.0f;
        if (i < cnt) {
            const float distanceNorm = c.trackLocation + ((P.lookAheadStep * (float)(i + 1)) / T.info.computedTrackLength) * driveDir;
            const V3 dir = track_direction_at_distance(T, distanceNorm);
            v = m_atan2(dot(cross(dir, curTrackDir), up), dot(curTrackDir, dir));
        }
        const float o = ex.peer(v);
        mine[0] = ex.half ? o : v; mine[1] = ex.half ? v : o;
    }
    for (int i = 0; i < cnt; ++i) c.lookAhead[i] = ex.get((i < 4) ? mine[0] : mine[1], i & 3);
}

/* ScoringSystem::step = computeDriftScore + computeAgentReward (ScoringSystem.cpp:114-330) */
PD_HD void post_scoring(const PdCarParams& P, const TrackDev& T, const Body& C, CarCtx& X, float dt) {
    CarS& c = X.c;
    /* ---- ScoringSystem::step (ScoringSystem.cpp:114-330) ---- */
    {
        /* validateDrift */
        bool bInvalid = true; int nDirty = 0;
        for (int w = 0; w < 4; ++w) { const int s = X.wl[w].surfaceId; if (s >= 0 && T.surfaces[s].dirtAdditiveK > 0.001f) nDirty++; }
        if (nDirty <= 2) { if ((c.speed * 3.6f) >= 20.0f) { if (!X.newDamage && c.currentGear) bInvalid = false; } }      /* bDamage: a damage zone rose this tick (ScoringSystem.cpp:359-368) */
        if (bInvalid) c.driftInvalid = 1;
        const float fBeta = fabsf(beta_rad(C));
        const float fSpeedKmh = c.speed * 3.6f;
        bool earlyOut = false;
        if (fSpeedKmh > 20.0f && fBeta > 0.13089749f) {
            const V3 v = irot(C.fr, C.v);
            if (!c.drifting) { c.lastDriftDirection = signf_(v.x); c.driftComboCounter = 1; c.driftInvalid = 0; c.instantDrift = 0.0f; }
            c.currentDriftAngle = fBeta - 0.13089749f;
            const float fSpeedMult = tclampf((fSpeedKmh - 20.0f) * 0.015384615f, 0.0f, 2.0f);
            c.currentSpeedMultiplier = fSpeedMult;
            int nDrifty = 0;
            for (int w = 0; w < 4; ++w) {
                const int s = X.wl[w].surfaceId;
                if (s >= 0 && fabsf(X.wl[w].angularVelocity) > 4.0 && fabsf(X.wl[w].slipRatio) > 0.8f && X.wl[w].load > 10.0 && T.surfaces[s].gripMod >= 0.9f) nDrifty++;
            }
            c.driftExtreme = nDrifty > 1;
            float fDelta = fSpeedMult * c.currentDriftAngle;
            if (c.driftExtreme) fDelta *= 2.0f;
            c.instantDriftDelta = fDelta; c.instantDrift += fDelta;
            if (fabsf(v.x) > 4.0f) {
                const float fDir = signf_(v.x);
                if (c.lastDriftDirection != fDir && fBeta > 0.26179498f) { c.instantDrift += 50.0f; c.driftComboCounter++; c.lastDriftDirection = fDir; }
            }
            c.drifting = 1; c.driftStraightTimer = 0.0f;
        }
        if (c.drifting) {
            if (fSpeedKmh > 20.0f && fBeta < 0.065448746f) c.driftStraightTimer += dt; else c.driftStraightTimer = 0.0f;
            if (c.driftInvalid) { c.currentDriftAngle = 0; c.currentSpeedMultiplier = 0; c.driftExtreme = 0; c.drifting = 0; c.instantDrift = 0; c.driftComboCounter = 0; earlyOut = true; }
            else if (c.driftStraightTimer > 1.0f) { c.driftComboCounter = 0; c.driftPoints += c.instantDrift; c.drifting = 0; c.instantDrift = 0.0f; }
        }
        if (!earlyOut && c.driftInvalid) { c.currentDriftAngle = 0; c.currentSpeedMultiplier = 0; c.driftExtreme = 0; c.drifting = 0; c.instantDrift = 0; c.driftComboCounter = 0; }
    }
    { /* computeAgentReward (ScoringSystem.cpp:129-248) */
        const float* SVv = P.scoring;
        float reward = 0.0f;
        const float curRpm = car_engine_rpm(c);
        const float maxRpm = (float)(int)(P.engine.limiter * P.engine.limiterMultiplier);
        if (c.oldPointId < c.nearestTrackPointId || (c.nearestTrackPointId == 0 && c.oldPointId != c.nearestTrackPointId)) { c.oldPointId = c.nearestTrackPointId; reward += SVv[PD_SV_TravelBonus]; }
        if (c.oldSplinePointId < c.splinePointId || (c.splinePointId == 0 && c.oldSplinePointId != c.splinePointId)) { c.oldSplinePointId = c.splinePointId; reward += SVv[PD_SV_TravelSplineBonus]; }
        reward += SVv[PD_SV_DriftBonus] * c.instantDriftDelta;
        reward += SVv[PD_SV_SpeedBonus] * linscalef(c.speed * 3.6f, SVv[PD_SV_MinBonusSpeed], SVv[PD_SV_MaxBonusSpeed], 0.0f, 1.0f);
        reward += SVv[PD_SV_ThrottleBonus] * linscalef(c.ctlGas, 0.0f, 1.0f, 0.0f, 1.0f);
        reward += SVv[PD_SV_EngineRpmBonus] * linscalef(curRpm, 0.0f, maxRpm, 0.0f, 1.0f);
        if (curRpm < SVv[PD_SV_StallRpm]) reward -= SVv[PD_SV_StallPenalty];
        if (c.isGearGrinding) reward -= SVv[PD_SV_GearGrindPenalty];
        if (P.nProbes > 0) {
            float closest = FLT_MAX;
            for (int i = 0; i < P.nProbes; ++i) { const float d = c.probes[i]; if (closest > d && d > 0.0f) closest = d; }
            if (closest < SVv[PD_SV_ApproachDistance]) reward -= SVv[PD_SV_ObstApproachPenalty] * (1.0f - linscalef(closest, SVv[PD_SV_CriticalDistance], SVv[PD_SV_ApproachDistance], 0.0f, 1.0f));
        }
        if (c.collisionFlag) reward -= SVv[PD_SV_CollisionPenalty];
        const int tp = c.nearestTrackPointId;
        if (tp >= 0 && tp < T.info.nFatPoints) {
            const PdFatPoint& pt = T.fat[tp];
            if (len(C.fr.p - v3(pt.center[0], pt.center[1], pt.center[2])) > T.info.computedTrackWidth * SVv[PD_SV_OutOfTrackThreshold]) { c.outOfTrackFlag = 1; reward -= SVv[PD_SV_OffTrackPenalty]; }
            const float x = c.bodyVsTrack;
            const float thresh = tclampf(SVv[PD_SV_DirectionThreshold], 0.1f, 1.0f);
            if (x > thresh) reward += SVv[PD_SV_DirectionBonus] * linscalef(x, thresh, 1.0f, 0.0f, 1.0f);
            else reward -= SVv[PD_SV_DirectionPenalty] * (1.0f - linscalef(x, -1.0f, thresh, 0.0f, 1.0f));
        }
        c.stepReward = reward; c.totalReward += reward;
    }
}

/* the tick */
template <int STRIDE, int STRIDE_D, int TOPO = 0, class SVX> PD_HDN void car_tick(const PdCarParams& P, const TrackDev& T, const SVX& sv, float dt, double physicsTime, float* scratch, float* scratchD, int collPre = -1, const float* cont = nullptr) {
    CarS cLocal; CarS* cp = &cLocal;
    if constexpr (sv_traits<SVX>::in_place) cp = car_in_place(sv); else load_car(sv, cLocal);
    CarCtx X(*cp); X.dt = dt; X.time = physicsTime;
    constexpr bool FDW = PD_TOPO_FRONT_DW(TOPO), RDW = PD_TOPO_REAR_DW(TOPO);
    Body bod[PD_NUM_BODIES]; V3 steerAnchor1[2], steerAnchor2[2];
    set_body_mass(bod, P, TOPO);
    PD_UNROLL
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body<TOPO>(i)) load_body(sv, i, bod[i]);
    CarS& c = X.c;
    Body& C = bod[PD_BODY_CHASSIS];

    /* ---------------- Simulator::stepCars: stepPreCacheValues + Car::step ---------------- */
    c.speed = len(C.v);
    c.collisionFlag = 0; c.outOfTrackFlag = 0;
    { /* Car.cpp:426-451 (car id 0): DBall ERP by speed; CFM = baseCFM = 1e-7 in both branches */
        const float fVelSq = sqlen(C.v);
        X.dballErp = (fVelSq >= 1.0f) ? 0.3f : 0.9f;
        X.dballCfm = (fVelSq >= 1.0f) ? susp_base_cfm<TOPO>(P) : 0.0000001f;
    }
    c.ctlSteer = tclampf(c.ctlSteer, -1.0f, 1.0f); c.ctlClutch = tclampf(c.ctlClutch, 0.0f, 1.0f); c.ctlBrake = tclampf(c.ctlBrake, 0.0f, 1.0f);
    c.ctlHandBrake = tclampf(c.ctlHandBrake, 0.0f, 1.0f); c.ctlGas = tclampf(c.ctlGas, 0.0f, 1.0f);
    {
        const float target = c.ctlSteer;
        if (c.smoothSteer) { const float diff = target - c.smoothSteerValue; c.smoothSteerValue += diff * P.scoring[PD_SV_SmoothSteerSpeed] * dt; c.ctlSteer = c.smoothSteerValue; }
        else c.smoothSteerValue = target;
    }
    { /* fuel (Car.cpp:476-489) */
        const float fRpmAbs = fabsf(engine_rpm(c));
        const double fNewFuel = c.fuel - (fRpmAbs * dt * c.gasUsage) * (tmaxf(0.0f, c.turboBoost) + 1.0) * P.fuelConsumptionK * 0.001 * P.fuelConsumptionRate;
        c.fuel = fNewFuel;
        if (fNewFuel > 0.0f) c.fuelPressure = 1.0f; else { c.fuel = 0; c.fuelPressure = 0; }
    }
    {
        float sig = (P.steerLock * c.ctlSteer) / P.steerRatio;
        if (!finitef(sig)) sig = 0;
        c.finalSteerAngleSignal = sig;
    }
    bool bAllTyresLoaded = true;
    for (int w = 0; w < 4; ++w) if (sv.f(PD_OFF_TYRE(w) + PD_TYRE_o_load) <= 0.0f) { bAllTyresLoaded = false; break; }
    autoclutch_step(P, X);
    {
        const float fAngVelSq = sqlen(C.w);
        if (c.speed >= 0.5f || fAngVelSq >= 1.0f) c.sleepingFrames = 0;
        else {
            if (bAllTyresLoaded && (c.ctlGas <= 0.01f || c.ctlClutch <= 0.01f || c.currentGear == 1)) c.sleepingFrames++; else c.sleepingFrames = 0;
            if (c.sleepingFrames > P.framesToSleep) { body_stop(C); body_stop(bod[PD_BODY_TANK]); }
        }
    }
    {
        const V3 vBodyVel = C.v;
        const V3 vAccel = (vBodyVel - v3(c.lastVelX, c.lastVelY, c.lastVelZ)) * (1.0f / dt) * 0.10197838f;
        c.lastVelX = vBodyVel.x; c.lastVelY = vBodyVel.y; c.lastVelZ = vBodyVel.z;
        const V3 g = irot(C.fr, vAccel);
        c.accGX = g.x; c.accGY = g.y; c.accGZ = g.z;
    }
    { /* stepThermalObjects (Car.cpp:624-634) + ThermalObject::step */
        const float fRpm = engine_rpm(c);
        float heat = 0;
        if (fRpm > (P.engine.minimum * 0.8f)) { const float fLimiter = (float)(int)(P.engine.limiter * P.engine.limiterMultiplier); heat += (((fRpm / fLimiter) * 20.0f) * c.ctlGas) + 85.0f; }
        const float fOneDivMass = 1.0f / P.waterTmass;
        const float fCool = 1.0f - (P.waterCoolSpeedK * c.speed);
        c.waterT += (((((fCool * P.ambientTemperature) - c.waterT) * fOneDivMass) * dt) * P.waterCoolFactor);
        if (heat != 0.0f) c.waterT += ((((heat - c.waterT) * fOneDivMass) * dt) * P.waterHeatFactor);
    }

    /* ---------------- stepComponents ---------------- */
    float brakeT[4], handT[4];
    brakes_step(P.brakes, c, sv, P.ambientTemperature, dt, brakeT, handT);
    float travel[4], dspeed[4];
    if constexpr (FDW) { dw_step(P.dw[0], C, bod[PD_BODY_HUB0], travel[0], dspeed[0]); dw_step(P.dw[1], C, bod[PD_BODY_HUB1], travel[1], dspeed[1]); }
    else { strut_step(P.strut[0], C, bod[PD_BODY_HUB0], travel[0], dspeed[0]); strut_step(P.strut[1], C, bod[PD_BODY_HUB1], travel[1], dspeed[1]); }
    if constexpr (RDW) { dw_step(P.dw[2], C, bod[PD_BODY_HUB2], travel[2], dspeed[2]); dw_step(P.dw[3], C, bod[PD_BODY_HUB3], travel[3], dspeed[3]); }
    else { axle_step(P.axle, C, bod[PD_BODY_AXLE], 0, travel[2], dspeed[2]); axle_step(P.axle, C, bod[PD_BODY_AXLE], 1, travel[3], dspeed[3]); }
    for (int w = 0; w < 4; ++w) {
        sv.f(PD_OFF_TYRE(w) + PD_TYRE_o_suspTravel, travel[w]); sv.f(PD_OFF_TYRE(w) + PD_TYRE_o_suspDamperSpeed, dspeed[w]);
    }
    for (int w = 0; w < 4; ++w) {
        if (w < 2) { Body& H = bod[PD_BODY_HUB0 + 2 * w]; const Frame hf = FDW ? dw_hub_frame(P.dw[w], H) : strut_hub_frame(P.strut[w], H); tyre_step(P, T, w, X, sv, H, hf, C, brakeT[w], handT[w], X.wl[w]); }
        else if constexpr (RDW) { Body& H = bod[PD_BODY_HUB2 + (w - 2)]; const Frame hf = dw_hub_frame(P.dw[w], H); tyre_step(P, T, w, X, sv, H, hf, C, brakeT[w], handT[w], X.wl[w]); }
        else { Body& A = bod[PD_BODY_AXLE]; const Frame hf = axle_hub_frame(P.axle, A, w - 2); tyre_step(P, T, w, X, sv, A, hf, C, brakeT[w], handT[w], X.wl[w]); }
    }
    /* heave springs (Car.cpp:655-659: after the tyres, only when their rate is non-zero) */
    if constexpr (FDW) { if (P.heave[0].present && P.heave[0].k != 0.0f) heave_step(P.heave[0], P.dw[0], P.dw[1], C, bod[PD_BODY_HUB0], bod[PD_BODY_HUB0].fr.p, bod[PD_BODY_HUB0].v, bod[PD_BODY_HUB1], bod[PD_BODY_HUB1].fr.p, bod[PD_BODY_HUB1].v, -1); }
    if constexpr (RDW) { if (P.heave[1].present && P.heave[1].k != 0.0f) heave_step(P.heave[1], P.dw[2], P.dw[3], C, bod[PD_BODY_HUB2], bod[PD_BODY_HUB2].fr.p, bod[PD_BODY_HUB2].v, bod[PD_BODY_HUB3], bod[PD_BODY_HUB3].fr.p, bod[PD_BODY_HUB3].v, -1); }
    aero_step(P, C);
    { /* SteeringSystem::step -> setSteerLengthOffset -> reseatDistanceJointLocal (incl. its local->world->local round trip) */
        const float steer = -c.finalSteerAngleSignal * P.steerLinearRatio;
        for (int s = 0; s < 2; ++s) {
            const float* refPoint = FDW ? P.dw[s].refPoint : P.strut[s].refPoint;
            const float* baseCarSteer = FDW ? P.dw[s].baseCarSteer : P.strut[s].baseCarSteer;
            const float* tyreSteer = FDW ? P.dw[s].tyreSteer : P.strut[s].tyreSteer;
            const float sx = signf_(refPoint[0]);
            const float offx = 0.0f + steer + (sx * (FDW ? P.dw[s].toeOutLinear : P.strut[s].toeOutLinear));
            const V3 carSteer = v3(baseCarSteer[0] + offx, baseCarSteer[1], baseCarSteer[2]);
            const Body& H = bod[PD_BODY_HUB0 + 2 * s];
            steerAnchor1[s] = to_local(C.fr, to_world(C.fr, carSteer));
            steerAnchor2[s] = to_local(H.fr, to_world(H.fr, v3(tyreSteer[0], tyreSteer[1], tyreSteer[2])));
        }
    }
    autoblip_step(P, X);
    autoshift_step(P, X);
    gearchanger_step(P, X);
    { const float fAxleTorq = drivetrain_step(P, X); if constexpr (!RDW) { add_rel_torque(C, v3(0, 0, fAxleTorq)); add_rel_torque(bod[PD_BODY_AXLE], v3(0, 0, -fAxleTorq)); } }   /* Drivetrain.cpp:547-553: only with a rigid rear axle */
    { /* driven wheels: angular velocity / lock state written by the drivetrain */
        const int dl = (P.drivetrain.tractionType == 1) ? 0 : 2;
        for (int w = dl; w < dl + 2; ++w) { sv.f(PD_OFF_TYRE(w) + PD_TYRE_o_angularVelocity, X.wl[w].angularVelocity); sv.i(PD_OFF_TYRE(w) + PD_TYRE_o_isLocked, X.wl[w].isLocked); }
    }
    arb_step(P.arbK[0], C, bod[PD_BODY_HUB0], bod[PD_BODY_HUB0].fr.p, bod[PD_BODY_HUB1], bod[PD_BODY_HUB1].fr.p);
    if constexpr (RDW) arb_step(P.arbK[1], C, bod[PD_BODY_HUB2], bod[PD_BODY_HUB2].fr.p, bod[PD_BODY_HUB3], bod[PD_BODY_HUB3].fr.p);
    else {
        Body& A = bod[PD_BODY_AXLE];
        const Frame f0 = axle_hub_frame(P.axle, A, 0), f1 = axle_hub_frame(P.axle, A, 1);
        arb_step(P.arbK[1], C, A, f0.p, A, f1.p);
    }

    /* ---------------- physics->step(dt): collisionStep (odd frames: car vs static meshes), then dWorldStep ---------------- */
    /* collPre: the answer of k_collide for this tick's start pose (0 / 1), or -1 = not computed: test here */
    if (c.physFrame & 1) { if (collPre >= 0 ? (collPre != 0) : car_collide(P, T, C, 0, 1)) c.collisionFlag = 1; }
    const bool freshContacts = (c.physFrame & 1) != 0;       /* contact joints made by this frame's collisionStep (odd) or left over from the last one (even) */
    c.physFrame++;
#if PD_SOLVER2
    float dmg[5] = {c.damageZone0, c.damageZone1, c.damageZone2, c.damageZone3, c.damageZone4};
    world_step2<TOPO>(P, bod, steerAnchor1, steerAnchor2, X.dballErp, X.dballCfm, dt, cont, freshContacts, c.lifeLeft, dmg);
    X.newDamage = fabsf(dmg[0] - c.damageZone0) > 0.001f || fabsf(dmg[1] - c.damageZone1) > 0.001f || fabsf(dmg[2] - c.damageZone2) > 0.001f || fabsf(dmg[3] - c.damageZone3) > 0.001f || fabsf(dmg[4] - c.damageZone4) > 0.001f;
    c.damageZone0 = dmg[0]; c.damageZone1 = dmg[1]; c.damageZone2 = dmg[2]; c.damageZone3 = dmg[3]; c.damageZone4 = dmg[4];
#else
    world_step<STRIDE, STRIDE_D>(P, bod, steerAnchor1, steerAnchor2, X.dballErp, X.dballCfm, dt, scratch, scratchD);
#endif

    /* ---------------- Car::postStep ---------------- */
    { /* updateTrackLocator (Car.cpp:717-771) */
        const V3 bodyPos = C.fr.p;
        const int nFat = T.info.nFatPoints;
        if (P.nProbes > 0 && nFat > 0) {
            /* Track::rayCastTrackBounds cache refresh (Track.cpp:505-510): only the first probe can trigger it */
            const V3 cache = v3(c.pointCacheX, c.pointCacheY, c.pointCacheZ);
            if (sqlen(cache - bodyPos) > 1.0f * 1.0f) { c.pointCacheX = bodyPos.x; c.pointCacheY = bodyPos.y; c.pointCacheZ = bodyPos.z; }
        }
        const V3 cachePos = v3(c.pointCacheX, c.pointCacheY, c.pointCacheZ);
        const float nearR = (P.nProbes > 0) ? P.probeLength[0] : T.info.hashCellSize;
        const float nearRSq = nearR * nearR;
        /* probe rays */
        float rax = bodyPos.x, raz = bodyPos.z;
        float rbx[PD_MAX_PROBES], rbz[PD_MAX_PROBES], best[PD_MAX_PROBES];
        for (int r = 0; r < P.nProbes; ++r) {
            const V3 rayStart = to_world(C.fr, v3(0, 0, 0));
            const V3 rayEndL = to_world(C.fr, v3(P.probeDir[r][0], P.probeDir[r][1], P.probeDir[r][2]) * P.probeLength[r]);
            const V3 dir = norm(rayEndL - rayStart);
            const V3 rayEnd = rayStart + dir * (P.probeLength[r] * 1.1f);
            rax = rayStart.x; raz = rayStart.z;
            rbx[r] = rayEnd.x; rbz[r] = rayEnd.z; best[r] = FLT_MAX;
        }
        float bestDistSq = FLT_MAX; int bestPoint = 0;
        bool needBrute = false;
        for (int r = 0; r < P.nProbes; ++r) if (!probe_walk(T, rax, raz, rbx[r], rbz[r], cachePos, nearRSq, best[r])) needBrute = true;
        const bool haveNearest = nearest_point_grid(T, bodyPos, cachePos, nearRSq, bestPoint);
        if (needBrute || !haveNearest) {
            /* exhaustive form of the reference (car far off the indexed area) */
            bestPoint = 0;
            for (int r = 0; r < P.nProbes; ++r) best[r] = FLT_MAX;
            for (int id = 0; id < nFat; ++id) {
                const PdFatPoint& f = T.fat[id];
                const V3 loc = v3(f.best[0], f.best[1], f.best[2]);
                if (!(sqlen(cachePos - loc) < nearRSq)) continue;          /* VertexHash::queryNeighbours filter */
                { const float dsq = sqlen(loc - bodyPos); if (bestDistSq > dsq) { bestDistSq = dsq; bestPoint = id; } }  /* getPointIdAtLocation */
                const PdFatPoint& g = T.fat[id + 1 < nFat ? id + 1 : 0];
                for (int r = 0; r < P.nProbes; ++r) {
                    float ix, iz;
                    if (line_intersection(rax, raz, rbx[r], rbz[r], f.left[0], f.left[2], g.left[0], g.left[2], ix, iz)) { const float dx = rax - ix, dz = raz - iz; best[r] = tminf(best[r], sqrtf(dx * dx + dz * dz)); }
                    if (line_intersection(rax, raz, rbx[r], rbz[r], f.right[0], f.right[2], g.right[0], g.right[2], ix, iz)) { const float dx = rax - ix, dz = raz - iz; best[r] = tminf(best[r], sqrtf(dx * dx + dz * dz)); }
                }
            }
        }
        for (int r = 0; r < P.nProbes; ++r) c.probes[r] = (best[r] != FLT_MAX) ? best[r] : P.probeLength[r];
        if (c.nearestTrackPointId != bestPoint) { c.oldTrackPointId = c.nearestTrackPointId; c.nearestTrackPointId = bestPoint; c.lastTrackPointTimestamp = (float)physicsTime; }
        c.oldTrackLocation = c.trackLocation; c.trackLocation = 0;
        if (bestPoint >= 0 && bestPoint < nFat) {
            if (nFat >= 5) { /* getDistanceAlongSplineAtLocation (Track.cpp:609-629) */
                int prevId = bestPoint - 1; if (prevId < 0) prevId = nFat - 1;
                int nextId = bestPoint + 1; if (nextId >= nFat) nextId = 0;
                int prevId2 = prevId - 1; if (prevId2 < 0) prevId2 = nFat - 1;
                int nextId2 = nextId + 1; if (nextId2 >= nFat) nextId2 = 0;
                int sid; float sdist;
                if (spline_nearest(T, bodyPos, prevId2 * T.info.interpolateStep, nextId2 * T.info.interpolateStep, sid, sdist)) {
                    c.splinePointId = sid; c.trackLocation = tclampf(sdist / T.info.computedTrackLength, 0.0f, 1.0f);
                }
            }
            const V3 bodyFrontDir = norm(C.fr.az);
            const V3 bodyVelDir = norm(C.v);
            const PdFatPoint& pt = T.fat[bestPoint];
            const V3 fwd = v3(pt.forwardDir[0], pt.forwardDir[1], pt.forwardDir[2]);
            c.bodyVsTrack = dot(bodyFrontDir, fwd);
            if ((c.speed * 3.6f) > 3.0f) c.velocityVsTrack = dot(bodyVelDir, fwd); else c.velocityVsTrack = 0.0f;
        }
    }
    post_lookahead(P, T, C, c);
    post_scoring(P, T, C, X, dt);
    c.episodeSteps++; c.thermalPrimed = 1;
    {
        int bad = 0;
        float acc = 0.0f;
        PD_UNROLL
        for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body<TOPO>(i)) acc += body_nonfinite_acc(bod[i]);
        bad = (acc == 0.0f) ? 0 : 1;
        if (bad) c.nanFlag = 1;
    }
    PD_UNROLL
    for (int i = 0; i < PD_NUM_BODIES; ++i) if (topo_has_body<TOPO>(i)) store_body(sv, i, bod[i]);
    if constexpr (!sv_traits<SVX>::in_place) store_car(sv, c);
}

/* observation vector of pyprojectd/projectd_env.py:237-275 */
template <class SVX> PD_HD void car_observe(const SVX& sv, float* obs /* 24, stride 1 */) {
    Body C; load_body(sv, PD_BODY_CHASSIS, C);
    const V3 lv = irot(C.fr, C.v), lw = irot(C.fr, C.w);
    obs[0] = lv.x; obs[1] = lv.y; obs[2] = lv.z; obs[3] = lw.x; obs[4] = lw.y; obs[5] = lw.z;
    for (int w = 0; w < 4; ++w) obs[6 + w] = sv.f(PD_OFF_TYRE(w) + PD_TYRE_o_ndSlip);
    obs[10] = sv.f(PD_OFF_CAR + PD_CAR_o_bodyVsTrack); obs[11] = sv.f(PD_OFF_CAR + PD_CAR_o_velocityVsTrack);
    for (int i = 0; i < 5; ++i) obs[12 + i] = sv.f(PD_OFF_LOOKAHEAD + i);
    for (int i = 0; i < 7; ++i) obs[17 + i] = sv.f(PD_OFF_PROBES + i);
}

/* the env's action mapping (pyprojectd/projectd_env.py:158-170): steer = a0, gas = linscale(a1, -1..1 -> min_gas..max_gas),
 * clutch = 1.0 when auto_clutch is off, requestedGearIndex = 2 when auto_shift is off, `smooth` = smooth_controls */
template <class SVX> PD_HD void env_apply_action(const SVX& sv, float a0, float a1, const PdEnvConfig& cfg) {
    const int o = PD_OFF_CAR;
    sv.f(o + PD_CAR_o_ctlSteer, a0); sv.f(o + PD_CAR_o_ctlClutch, cfg.clutch); sv.f(o + PD_CAR_o_ctlBrake, 0.0f); sv.f(o + PD_CAR_o_ctlHandBrake, 0.0f);
    sv.f(o + PD_CAR_o_ctlGas, linscalef(a1, -1.0f, 1.0f, cfg.min_gas, cfg.max_gas));
    sv.i(o + PD_CAR_o_ctlRequestedGear, cfg.requested_gear); sv.i(o + PD_CAR_o_ctlGearUp, 0); sv.i(o + PD_CAR_o_ctlGearDn, 0); sv.i(o + PD_CAR_o_smoothSteer, cfg.smooth_controls ? 1 : 0);
}

/* ProjectDEnv.step tail (projectd_env.py:178-212): reward with the termination penalties and the PD_DONE_* causes
 * that depend on the car alone (the low-reward cut needs the env's running return and is applied by the caller) */
template <class SVX> PD_HD void env_reward_done(const SVX& sv, double timeAfter, const PdEnvConfig& cfg, float& reward, int& done) {
    const int o = PD_OFF_CAR;
    float r = sv.f(o + PD_CAR_o_stepReward);
    int d = 0;
    if (cfg.terminate_on_hit && sv.i(o + PD_CAR_o_collisionFlag)) { r -= cfg.terminate_hit_penalty; d |= 1 /* PD_DONE_COLLISION */; }
    if (cfg.terminate_off_track && sv.i(o + PD_CAR_o_outOfTrackFlag)) { r -= cfg.terminate_off_track_penalty; d |= 2 /* PD_DONE_OFFTRACK */; }
    /* the env compares lastTrackPointTimestamp + stuck_timeout against the state's timestamp */
    if (cfg.terminate_when_stuck && sv.f(o + PD_CAR_o_lastTrackPointTimestamp) + cfg.stuck_timeout < (float)timeAfter) { r -= cfg.terminate_stuck_penalty; d |= 4 /* PD_DONE_STUCK */; }
    if (sv.i(o + PD_CAR_o_nanFlag)) d |= 16 /* PD_DONE_NAN */;
    reward = r; done = d;
}

} // namespace pd
