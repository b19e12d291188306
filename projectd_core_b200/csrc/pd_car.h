/*
 * pd_car.h -- per-car components of Car::step as device functions.
 *
 * One thread advances one car.  Evaluation order, the order in which forces are accumulated on the
 * bodies, and the float/double split all follow the reference (cited per function, paths relative to
 * src/ProjectD/): several components read state written earlier in the same tick, so order is semantics.
 */
#pragma once
#include "pd_track.h"

namespace pd {

/* what later components need to know about a wheel after Tyre::step */
struct WheelLink {
    float load, feedbackTorque, angularVelocity, brakeTorque, handBrakeTorque, ndSlip, slipRatio;
    int isLocked, surfaceId;
};

/* car-level (non-body) context of one tick; every lane of a car's quad holds an identical copy */
struct CarCtx {
    CarS& c;              /* a local copy (tiled state) or the record's car part itself (stride-1 view, in place) */
    WheelLink wl[PD_NUM_WHEELS];
    float dt;
    double time;          /* Simulator::physicsTime */
    float dballErp, dballCfm;
#if defined(PD_PHASE_CLOCKS)
    long long* ph = nullptr;   /* profiling build: phase time stamps of this warp (written by its lane 0) */
#endif
    bool newDamage = false;      /* a damage zone rose during this tick's collision callbacks */
    PD_HD explicit CarCtx(CarS& cc) : c(cc) {}
};
#if defined(PD_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
#define PD_PHASE(X, k) do { if ((X).ph) (X).ph[k] = clock64(); } while (0)
#else
#define PD_PHASE(X, k) do { } while (0)
#endif

PD_HD float engine_rpm(const CarS& c) { return (float)((c.engineVel * 0.15915507) * 60.0); }   /* Drivetrain::getEngineRPM */
PD_HD float car_engine_rpm(const CarS& c) { return ((float)c.engineVel * 0.15915507f * 60.0f); } /* Car::getEngineRpm */

/* ---- hub frames (ISuspension::getHubWorldMatrix) ---- */
PD_HD Frame strut_hub_frame(const PdStrut& P, const Body& hub) { /* SuspensionStrut.cpp:360-366 */
    Frame f; const float s = m_sin(P.staticCamber), c = m_cos(P.staticCamber);
    f.ax = v3(c * hub.fr.ax.x + s * hub.fr.ay.x, c * hub.fr.ax.y + s * hub.fr.ay.y, c * hub.fr.ax.z + s * hub.fr.ay.z);
    f.ay = v3(-s * hub.fr.ax.x + c * hub.fr.ay.x, -s * hub.fr.ax.y + c * hub.fr.ay.y, -s * hub.fr.ax.z + c * hub.fr.ay.z);
    f.az = hub.fr.az; f.p = hub.fr.p;
    return f;
}
PD_HD Frame dw_hub_frame(const PdDW& P, const Body& hub) { /* SuspensionDW.cpp:333-336: rotate(hub world matrix, z, staticCamber) */
    Frame f; const float s = m_sin(P.staticCamber), c = m_cos(P.staticCamber);
    f.ax = v3(c * hub.fr.ax.x + s * hub.fr.ay.x, c * hub.fr.ax.y + s * hub.fr.ay.y, c * hub.fr.ax.z + s * hub.fr.ay.z);
    f.ay = v3(-s * hub.fr.ax.x + c * hub.fr.ay.x, -s * hub.fr.ax.y + c * hub.fr.ay.y, -s * hub.fr.ax.z + c * hub.fr.ay.z);
    f.az = hub.fr.az; f.p = hub.fr.p;
    return f;
}
PD_HD Frame axle_hub_frame(const PdAxle& P, const Body& axle, int side) { /* SuspensionAxle.cpp:224-232 */
    Frame f = axle.fr; const float t = side ? -P.track : P.track;
    f.p.x += f.ax.x * t; f.p.y += f.ax.y * t; f.p.z += f.ax.z * t;
    return f;
}

PD_HD float damper_force(const PdDamper& d, float speed) { /* Damper.cpp:11-29 */
    float f;
    if (speed <= 0.0f) {
        if (fabsf(speed) <= d.fastThresholdRebound) f = -(speed * d.reboundSlow);
        else f = (d.fastThresholdRebound * d.reboundSlow) - ((d.fastThresholdRebound + speed) * d.reboundFast);
    } else {
        if (speed <= d.fastThresholdBump) f = -(speed * d.bumpSlow);
        else f = -(((speed - d.fastThresholdBump) * d.bumpFast) + (d.fastThresholdBump * d.bumpSlow));
    }
    return f;
}

/* SuspensionStrut::step (SuspensionStrut.cpp:230-290) */
PD_HD void strut_step(const PdStrut& P, Body& C, Body& H, float& travelOut, float& damperSpeedOut) {
    const V3 vCarStrut = to_world(C.fr, v3(P.carStrut[0], P.carStrut[1], P.carStrut[2]));
    const V3 vTyreStrut = to_world(H.fr, v3(P.tyreStrut[0], P.tyreStrut[1], P.tyreStrut[2]));
    V3 vDelta = vTyreStrut - vCarStrut;
    const float fDeltaLen = len(vDelta);
    vDelta = norm_l(vDelta, fDeltaLen);
    const float fDefaultLength = P.strutBaseLength + P.rodLength;
    const float fTravel = fDefaultLength - fDeltaLen;
    travelOut = fTravel;
    float fForce = ((fTravel * P.progressiveK) + P.k) * fTravel;
    if (fForce < 0) fForce = 0;
    if (P.packerRange != 0.0f && fTravel > P.packerRange) fForce += ((fTravel - P.packerRange) * P.bumpStopRate);
    if (fForce > 0) {
        const V3 vForce = vDelta * fForce;
        add_force_at_pos(H, vForce, vTyreStrut);
        add_force_at_pos(C, vForce * -1.0f, vCarStrut);
    }
    const V3 vHubWorld = H.fr.p;
    const V3 vHubLocal = to_local(C.fr, vHubWorld);
    const float fHubDelta = vHubLocal.y - P.refPoint[1];
    if (fHubDelta > P.bumpStopUp) {
        fForce = (fHubDelta - P.bumpStopUp) * 500000.0f;
        add_force_at_pos(H, C.fr.ay * -fForce, H.fr.p);
        add_rel_force_at_rel_pos(C, v3(0, fForce, 0), vHubLocal);
    }
    if (fHubDelta < P.bumpStopDn) {
        fForce = (fHubDelta - P.bumpStopDn) * 500000.0f;
        add_force_at_pos(H, C.fr.ay * -fForce, H.fr.p);
        add_rel_force_at_rel_pos(C, v3(0, fForce, 0), vHubLocal);
    }
    const V3 vTyreStrutVel = body_rel_point_vel(H, v3(P.tyreStrut[0], P.tyreStrut[1], P.tyreStrut[2]));
    const V3 vCarStrutVel = body_rel_point_vel(C, v3(P.carStrut[0], P.carStrut[1], P.carStrut[2]));
    const V3 vDamperDelta = vTyreStrutVel - vCarStrutVel;
    const float fDamperSpeed = dot(vDamperDelta, vDelta);
    damperSpeedOut = fDamperSpeed;
    const float fDamperForce = damper_force(P.damper, fDamperSpeed);
    const V3 vDamperForce = vDelta * fDamperForce;
    add_force_at_pos(H, vDamperForce, vTyreStrut);
    add_force_at_pos(C, vDamperForce * -1.0f, vCarStrut);
}

/* SuspensionAxle::step (SuspensionAxle.cpp:120-185), side 0 = Left, 1 = Right */
PD_HD void axle_step(const PdAxle& P, Body& C, Body& A, int side, float& travelOut, float& damperSpeedOut) {
    const float fSideSign = side ? -1.0f : 1.0f;
    const V3 vAxleWorld = (A.fr.ax * (fSideSign * P.track * P.attachRelativePos)) + A.fr.p;
    const V3 vAxleLocal = to_local(C.fr, vAxleWorld);
    const float t = side ? -P.track : P.track;
    V3 vBase = v3(t, P.axleBasePos[1], P.axleBasePos[2]);
    vBase.x *= P.attachRelativePos;
    vBase.y += 0.2f;
    const V3 vBaseWorld = to_world(C.fr, vBase);
    V3 vDelta = vBaseWorld - vAxleWorld;
    const float fDeltaLen = len(vDelta);
    vDelta = norm_l(vDelta, fDeltaLen);
    const float fTravel = (0.2f - fDeltaLen) + P.rodLength;
    travelOut = fTravel;
    float fForce = -(((fTravel * P.progressiveK) + P.k) * fTravel);
    if (fForce < 0.0f) {
        add_force_at_pos(A, vDelta * fForce, vAxleWorld);
        add_force_at_pos(C, vDelta * -fForce, vBaseWorld);
    }
    if (P.leafSpringKx != 0.0f) {
        fForce = (vAxleLocal.x - (P.attachRelativePos * t)) * P.leafSpringKx;
        add_force_at_pos(A, C.fr.ax * -fForce, vAxleWorld);
        add_rel_force_at_rel_pos(C, v3(fForce, 0.0f, 0.0f), vAxleLocal);
    }
    const float fRefY = vAxleLocal.y - P.referenceY;
    if (P.bumpStopUp != 0.0f && fRefY > P.bumpStopUp && 0.0f != P.k) {
        fForce = (fRefY - P.bumpStopUp) * 500000.0f;
        add_force_at_pos(A, C.fr.ay * -fForce, vAxleWorld);
        add_rel_force_at_rel_pos(C, v3(0.0f, fForce, 0.0f), vAxleLocal);
    }
    if (P.bumpStopDn != 0.0f && fRefY < P.bumpStopDn && 0.0f != P.k) {
        fForce = (fRefY - P.bumpStopDn) * 500000.0f;
        add_force_at_pos(A, C.fr.ay * -fForce, vAxleWorld);
        add_rel_force_at_rel_pos(C, v3(0.0f, fForce, 0.0f), vAxleLocal);
    }
    const V3 vPointVel = body_point_vel(C, vBaseWorld);
    const V3 vDeltaVel = body_rel_point_vel(A, v3(t, 0, 0)) - vPointVel;
    const float fDamperSpeed = dot(vDeltaVel, vDelta);
    damperSpeedOut = fDamperSpeed;
    const float fDamperForce = damper_force(P.damper, fDamperSpeed);
    const V3 vForce = vDelta * fDamperForce;
    add_force_at_pos(A, vForce, vAxleWorld);
    add_force_at_pos(C, vForce * -1.0f, vBaseWorld);
}

/* SuspensionDW::step (SuspensionDW.cpp:202-272), passive branch (useActiveActuator is never set) */
PD_HD void dw_step(const PdDW& P, Body& C, Body& H, float& travelOut, float& damperSpeedOut) {
    const V3 vBodyM2 = C.fr.ay;
    const V3 vRef = v3(P.refPoint[0], P.refPoint[1], P.refPoint[2]);
    const V3 vHubWorldPos = H.fr.p;
    const V3 vHubLocalPos = to_local(C.fr, vHubWorldPos);
    const float fHubDeltaY = vHubLocalPos.y - P.refPoint[1];
    const float fTravel = fHubDeltaY + P.rodLength;
    travelOut = fTravel;
    float fForce = ((fTravel * P.progressiveK) + P.k) * fTravel;
    if (P.multilink) { /* SuspensionML::step (SuspensionML.cpp:105-137): plain packer, the force acts whatever its sign, no bump stops */
        if (P.packerRange != 0.0f && fTravel > P.packerRange && P.k != 0.0f) fForce += ((fTravel - P.packerRange) * P.bumpStopRate);
        add_force_at_pos(H, vBodyM2 * -fForce, vHubWorldPos);
        add_rel_force_at_rel_pos(C, v3(0.0f, fForce, 0.0f), vRef);
    } else {
    if (P.packerRange != 0.0f && fTravel > P.packerRange && P.k != 0.0f)
        fForce += (((fTravel - P.packerRange) * P.bumpStopProgressive) + P.bumpStopRate) * (fTravel - P.packerRange);
    if (fForce > 0.0f) {
        add_force_at_pos(H, vBodyM2 * -fForce, vHubWorldPos);
        add_rel_force_at_rel_pos(C, v3(0, fForce, 0), vRef);
    }
    }
    const V3 vDeltaVel = H.v - body_rel_point_vel(C, vRef);
    const float fDamperSpeed = dot(vDeltaVel, vBodyM2);
    damperSpeedOut = fDamperSpeed;
    const float fDamperForce = damper_force(P.damper, fDamperSpeed);
    const V3 vForce = vBodyM2 * fDamperForce;
    add_force_at_pos(H, vForce, vHubWorldPos);
    add_force_at_rel_pos(C, vForce * -1.0f, vRef);
    if (P.multilink) return;
    if (P.bumpStopUp != 0.0f && fHubDeltaY > P.bumpStopUp && 0.0f != P.k) {
        const float f = (((fHubDeltaY - P.bumpStopUp) * P.bumpStopProgressive) + P.bumpStopRate) * (fHubDeltaY - P.bumpStopUp);
        add_force_at_pos(H, vBodyM2 * -f, vHubWorldPos);
        add_rel_force_at_rel_pos(C, v3(0, f, 0), vHubLocalPos);
    }
    if (P.bumpStopDn != 0.0f && fHubDeltaY < P.bumpStopDn && 0.0f != P.k) {
        const float f = (((fHubDeltaY - P.bumpStopDn) * P.bumpStopProgressive) + P.bumpStopRate) * (fHubDeltaY - P.bumpStopDn);
        add_force_at_pos(H, vBodyM2 * -f, vHubWorldPos);
        add_rel_force_at_rel_pos(C, v3(0, f, 0), vHubLocalPos);
    }
}

/* HeaveSpring::step (HeaveSpring.cpp:56-149) between the hubs H0 (left) / H1 (right) of a double-wishbone axle.  `mine`: 0 / 1 = apply only what
 * acts on hub 0 / hub 1 and the chassis share at its reference point (one lane per corner), -1 = everything (one thread per car). */
PD_HD void heave_step(const PdHeave& Hv, const PdDW& D0, const PdDW& D1, Body& C, Body& H0, V3 hubPos0, V3 hubVel0, Body& H1, V3 hubPos1, V3 hubVel1, int mine) {
    const V3 vM2 = C.fr.ay;
    const V3 vRefPoint0 = v3(D0.refPoint[0], D0.refPoint[1], D0.refPoint[2]), vRefPoint1 = v3(D1.refPoint[0], D1.refPoint[1], D1.refPoint[2]);
    const V3 vHubLoc0 = to_local(C.fr, hubPos0), vHubLoc1 = to_local(C.fr, hubPos1);
    float rodLength = Hv.rodLength;
    if (D0.k != 0.0f || D1.k != 0.0f) rodLength = (D1.rodLength + D0.rodLength) * 0.5f;
    const float fAvgY = (vHubLoc0.y + vHubLoc1.y) * 0.5f;
    const float fTravel = (fAvgY - D0.refPoint[1]) + rodLength;
    auto apply = [&](V3 hubForce, V3 bodyLocalForce) {
        if (mine != 1) { add_force_at_pos(H0, hubForce, hubPos0); add_rel_force_at_rel_pos(C, bodyLocalForce, vRefPoint0); }
        if (mine != 0) { add_force_at_pos(H1, hubForce, hubPos1); add_rel_force_at_rel_pos(C, bodyLocalForce, vRefPoint1); }
    };
    float v12 = ((fTravel * Hv.progressiveK) + Hv.k) * fTravel;
    if (Hv.packerRange != 0.0f && fTravel > Hv.packerRange) v12 += ((fTravel - Hv.packerRange) * Hv.bumpStopRate);
    apply(vM2 * -v12, v3(0.0f, v12, 0.0f));
    const float fDeltaY0 = fAvgY - D0.refPoint[1];
    if (Hv.bumpStopUp != 0.0f && fDeltaY0 > Hv.bumpStopUp) { const float f = (fDeltaY0 - Hv.bumpStopUp) * 500000.0f; apply(vM2 * -f, v3(0.0f, f, 0.0f)); }
    if (Hv.bumpStopDn != 0.0f && fDeltaY0 < Hv.bumpStopDn) { const float f = (fDeltaY0 - Hv.bumpStopDn) * 500000.0f; apply(vM2 * -f, v3(0.0f, f, 0.0f)); }
    const V3 vHubVel = (hubVel0 + hubVel1) * 0.5f;
    const V3 vLpv = (body_rel_point_vel(C, vRefPoint0) + body_rel_point_vel(C, vRefPoint1)) * 0.5f;
    const float fDamperForce = damper_force(Hv.damper, dot(vHubVel - vLpv, vM2));
    V3 vForce = vM2 * fDamperForce;
    /* the reference hands the damper's chassis share to addLocalForceAtLocalPos although it is a world vector (HeaveSpring.cpp:145-148): kept as is */
    apply(vForce, vForce * -1.0f);
}

/* AntirollBar::step (AntirollBar.cpp:19-47) */
PD_HD void arb_step(float k, Body& C, Body& H0, V3 hubWorld0, Body& H1, V3 hubWorld1) {
    if (k > 0.0f) {
        const V3 vHubLoc0 = to_local(C.fr, hubWorld0);
        const V3 vHubLoc1 = to_local(C.fr, hubWorld1);
        const float fDelta = vHubLoc1.y - vHubLoc0.y;
        const float fDeltaK = fDelta * k;
        const V3 vForce = norm(C.fr.ay) * fDeltaK;
        add_force_at_pos(H0, vForce, hubWorld0);
        add_force_at_pos(H1, vForce * -1.0f, hubWorld1);
        add_rel_force_at_rel_pos(C, v3(0, -fDeltaK, 0), vHubLoc0);
        add_rel_force_at_rel_pos(C, v3(0, fDeltaK, 0), vHubLoc1);
    }
}

/* BrakeSystem::step (BrakeSystem.cpp:82-168): front share (fixed or EBBMode::Internal), torques, disc temperatures; the dynamic
 * controllers (ctrl_ebb.ini, steer_brake_controller.ini) are rejected by the loader.  sv: the tyres' state of the last tick */
template <class SVX> PD_HD void brakes_step(const PdBrakes& P, CarS& c, const SVX& sv, float ambient, float dt, float* brakeTorque, float* handBrakeTorque) {
    float fFrontBias = P.frontBias;
    if (P.ebbInternal) {
        const float fLoadFront = sv.f(PD_OFF_TYRE(1) + PD_TYRE_o_load) + sv.f(PD_OFF_TYRE(0) + PD_TYRE_o_load);
        const float fLoadAWD = (sv.f(PD_OFF_TYRE(3) + PD_TYRE_o_load) + sv.f(PD_OFF_TYRE(2) + PD_TYRE_o_load)) + fLoadFront;
        const bool bFlag = fLoadAWD != 0.0f && (c.speed * 3.6f) > 10.0f;
        fFrontBias = bFlag ? tclampf(((fLoadFront / fLoadAWD) * P.ebbFrontMultiplier), 0.0f, 1.0f) : P.frontBias;
    }
    fFrontBias = tclampf(fFrontBias, P.biasMin, P.biasMax);
    const float fBrakeInput = tmaxf(c.ctlBrake, 0.0f /* brakeOverride */);
    const float fBrakeTorq = (P.brakePower * P.brakePowerMultiplier) * fBrakeInput;
    brakeTorque[0] = fBrakeTorq * fFrontBias;
    brakeTorque[1] = fBrakeTorq * fFrontBias;
    float fRear = ((1.0f - fFrontBias) * fBrakeTorq) - 0.0f /* rearCorrectionTorque */;
    if (fRear < 0.0f) fRear = 0;
    brakeTorque[2] = fRear; brakeTorque[3] = fRear;
    handBrakeTorque[0] = 0; handBrakeTorque[1] = 0;
    handBrakeTorque[2] = c.ctlHandBrake * P.handBrakeTorque;
    handBrakeTorque[3] = c.ctlHandBrake * P.handBrakeTorque;
    if (P.hasTemps) { /* BrakeSystem::stepTemps (BrakeSystem.cpp:151-168) */
        const float fSpeed = c.speed * 3.6f;
        float* T[4] = {&c.brakeDiscT0, &c.brakeDiscT1, &c.brakeDiscT2, &c.brakeDiscT3};
        for (int i = 0; i < 4; ++i) {
            const PdBrakeDisc& D = P.disc[i];
            float t = *T[i];
            brakeTorque[i] = curve_value(D.perfCurve, t) * brakeTorque[i];
            const float fCool = ((fSpeed * D.coolSpeedFactor) + 1.0f) * D.coolTransfer;
            t += (((ambient - t) * fCool) * dt);
            t += (((fabsf(sv.f(PD_OFF_TYRE(i) + PD_TYRE_o_angularVelocity)) * (brakeTorque[i] * D.torqueK)) * 0.001f) * dt);
            *T[i] = t;
        }
    }
}

/* ================================ tyre ================================ */

/* SCTM::getPureFY (TyreModel.cpp:149-161) */
PD_HD float sctm_pure_fy(const PdTyre& P, float cf, float slip) {
    const float v5 = (cf * 2.0f) * 0.0064f;
    const float v6 = 1.0f / (v5 / 3.0f);
    float fy;
    if (v6 < slip) fy = ((1.0f / (((slip - v6) * P.falloffSpeed) + 1.0f)) * (1.0f - P.asy)) + P.asy;
    else fy = (((1.0f - (slip / v6)) * (1.0f - (slip / v6))) * (v5 * slip)) + ((3.0f - ((slip / v6) * 2.0f)) * ((slip / v6) * (slip / v6)));
    return fy;
}
PD_HD float sctm_static_dy(const PdTyre& P, float load) { if (load != 0.0f) return (m_pow(load, P.lsExpY) * P.lsMultY) / load; return 0; }
PD_HD float sctm_static_dx(const PdTyre& P, float load) { if (load != 0.0) return (m_pow(load, P.sctmLsExpX) * P.sctmLsMultX) / load; return 0; }

struct TmIn { float load, slipAngleRAD, slipRatio, camberRAD, speed, u, cpLength, grain, blister, pressureRatio; };
struct TmOut { float Fy, Fx, Mz, trail, ndSlip, Dy, Dx; };

/* SCTM::solve (TyreModel.cpp:11-119); load/camber LUT variants are rejected by the loader */
PD_HD TmOut sctm_solve(const PdTyre& P, const TmIn& tmi) {
    TmOut tmo; tmo.Fy = 0; tmo.Fx = 0; tmo.Mz = 0; tmo.trail = 0; tmo.ndSlip = 0; tmo.Dy = 0; tmo.Dx = 0;
    if (tmi.load <= 0.0f || (tmi.slipAngleRAD == 0.0f && tmi.slipRatio == 0.0f && tmi.camberRAD == 0.0f)) return tmo;
    const float fSlipAngle = tmi.slipAngleRAD;
    const float fUnk1 = (m_sin(tmi.camberRAD) * P.camberGain) + fSlipAngle;
    const float fUnk1Tan = m_tan(fUnk1);
    const float fSlipAngleSin = m_sin(fSlipAngle);
    const float fBlister1 = tclampf(tmi.blister * 0.01f, 0.0f, 1.0f);
    const float fBlister2 = (fBlister1 * 0.2f) + 1.0f;
    const float fStaticDy = sctm_static_dy(P, tmi.load);
    const float fStaticDx = sctm_static_dx(P, tmi.load);
    float fUDy = tmi.u * fStaticDy / fBlister2;
    float fUDx = tmi.u * fStaticDx / fBlister2;
    if (tmi.slipRatio < 0.0f) fUDx = fUDx * P.brakeDXMod;
    const float fCamberRad = tmi.camberRAD;
    float fCamberRadTmp = fabsf(fCamberRad);
    if ((fCamberRad < 0.0f || fUnk1 < 0.0f) && (fCamberRad > 0.0f || fUnk1 > 0.0f)) fCamberRadTmp = -fCamberRadTmp;
    fCamberRadTmp = -fCamberRadTmp;
    {
        float fCamberUnk = (fCamberRadTmp * P.dcamber0) - ((fCamberRadTmp * fCamberRadTmp) * P.dcamber1);
        if (fCamberUnk <= -1.0f) fCamberUnk = -0.8999999f;
        fUDy += (((fUDy / (fCamberUnk + 1.0f)) - fUDy) * P.dCamberBlend);
    }
    const float fSlipRatio = tmi.slipRatio;
    const float fSlipAngleCos = m_cos(tmi.slipAngleRAD);
    const float fSlipRatioClamped = (fSlipRatio > -0.9999999f ? fSlipRatio : -0.9999999f);
    const float fSpeed = tmi.speed;
    const float a = fSpeed * fSlipAngleSin;
    const float b = (fSpeed * fSlipRatio) * fSlipAngleCos;
    const float fUnk2 = sqrtf((a * a) + (b * b));
    const float fUnk2Scaled = fUnk2 * P.speedSensitivity;
    const float fDy = fUDy / (fUnk2Scaled + 1.0f);
    const float fDx = fUDx / (fUnk2Scaled + 1.0f);
    const float fLoadSubFz0 = tmi.load - P.sctmFz0;
    const float fCF = ((((1.0f / ((((fLoadSubFz0 / P.sctmFz0) * (P.maxSlip1 - P.maxSlip0)) + P.maxSlip0) * (((tmi.u - 1.0f) * 0.75f) + 1.0f))) * 3.0f) * 78.125f) / ((tmi.grain * 0.01f) + 1.0f)) * ((P.pressureCfGain * tmi.pressureRatio) + 1.0f);
    const float fUnk3 = fSlipRatio / (fSlipRatioClamped + 1.0f);
    const float fUnk4 = fUnk1Tan / (fSlipRatioClamped + 1.0f);
    float fSlip;
    const float fCombFactor = P.combinedFactor;
    if (fCombFactor <= 0.0f || fCombFactor == 2.0f) fSlip = sqrtf((fUnk4 * fUnk4) + (fUnk3 * fUnk3));
    else { const float c34 = m_pow(fabsf(fUnk4), fCombFactor) + m_pow(fabsf(fUnk3), fCombFactor); fSlip = m_pow(c34, 1.0f / fCombFactor); }
    const float fPureFyDx = sctm_pure_fy(P, fCF * P.cfXmult, fSlip) * fDx;
    const float fPureFyDy = sctm_pure_fy(P, fCF, fSlip);
    tmo.Fy = ((fPureFyDy * fDy) * (fUnk4 / fSlip)) * tmi.load;
    tmo.Fx = ((fUnk3 / fSlip) * fPureFyDx) * tmi.load;
    const float fNdSlip = fSlip / (1.0f / (((fCF * 2.0f) * 0.0064f) / 3.0f));
    const float fUnk5 = tclampf((1.0f - (fNdSlip * 0.8f)), 0.0f, 1.0f);
    const float fUnk6 = (((((3.0f - (fUnk5 * 2.0f)) * (fUnk5 * fUnk5)) * 1.1f) - 0.1f) * tmi.cpLength) * 0.12f;
    tmo.Mz = -(fUnk6 * tmo.Fy);
    tmo.trail = fUnk6 * tclampf(tmi.speed, 0.0f, 1.0f);
    tmo.ndSlip = fNdSlip; tmo.Dy = fDy; tmo.Dx = fDx;
    return tmo;
}

/* where the 12 x 3 thermal grid of a tyre lives while the tick works on it: in the TyreS copy (or the in-place record), or --
 * thread-per-car kernel on the tiled global state, sv_traits::grid_in_place -- in the state words themselves.  A warp's accesses
 * to one patch word are one 128-byte line either way; the local copy cost a local store + load per patch on the way in and again
 * on the way out (profiles/r02_local_memory_attribution.md: 16 % of the kernel's local-memory sectors). */
struct GridLocal { float* T; PD_HD float get(int p) const { return T[p]; } PD_HD void set(int p, float v) const { T[p] = v; } };
template <class SVX> struct GridSV { const SVX& sv; int base; PD_HD float get(int p) const { return sv.f(base + p); } PD_HD void set(int p, float v) const { sv.f(base + p, v); } };
/* the grid words of a batch larger than L2 come from HBM; the sweep is a rolled loop with a loop-carried dependency, so its first touch of every line would be
 * exposed latency: ask for the 36 lines (one 128-byte line per word and warp) when the tyre's step begins, a ray cast and the force model ahead of the sweep */
#ifndef PD_PUNCTURE_FUSED
#define PD_PUNCTURE_FUSED 1      /* in-place grid: stepPuncture's stripe sums come out of the thermal sweep (0: read the grid on the spot, as the other layouts do).
                                  * A/B on one B200, M car-ticks/s, fused / on the spot (both with the L1 prefetch): 65536 envs 87.3 / 86.0, 32768 envs (255-register instance) 64.2 / 65.9 */
#endif
#ifndef PD_GRID_PREFETCH
#define PD_GRID_PREFETCH 1
#endif
/* experiment knob PD_TYRE_PREFETCH_L2: while tyre w is stepped, ask L2 for the record of tyre w + 1 (90 words = 90 lines per warp).
 * Measured at 65536 envs on B200: 85.7 M car-ticks/s with, 87.0 M without (A/B on one box, twice each): off. */
#ifndef PD_TYRE_PREFETCH_L2
#define PD_TYRE_PREFETCH_L2 0
#endif
template <class SVX> PD_HD void tyre_prefetch_l2(const SVX& sv, int w) {
#if defined(__CUDA_ARCH__) && PD_TYRE_PREFETCH_L2
    PD_UNROLL
    for (int p = 0; p < PD_TYRE_WORDS; ++p) asm volatile("prefetch.global.L2 [%0];" ::"l"(sv.s + (PD_OFF_TYRE(w) + p) * SVX::stride));
#else
    (void)sv; (void)w;
#endif
}
template <class SVX> PD_HD void grid_prefetch(const SVX& sv, int base) {
#if defined(__CUDA_ARCH__) && PD_GRID_PREFETCH
    PD_UNROLL
    for (int p = 0; p < PD_THERMAL_PATCHES; ++p) asm volatile("prefetch.global.L1 [%0];" ::"l"(sv.s + (base + p) * SVX::stride));
#else
    (void)sv; (void)base;
#endif
}

/* thermal grid neighbours in the reference's connection order (TyreThermalModel.cpp:28-58, buildTyre):
 * patch index p = element + stripe * 12 */
template <class GRID> PD_HD void thermal_step(const PdTyre& P, TyreS& t, const GRID g, float inBase, int inElem, float in0, float in1, float in2, float coreTInput, float dt, float angularSpeed, float camberRAD, float ambient, float carSpeed, float* preSum = nullptr) {
    /* TyreThermalModel::step (TyreThermalModel.cpp:60-110).  preSum (optional, 3 floats): per stripe, the sum of the patch temperatures BEFORE this step
     * in element order -- what stepPuncture averages -- so that a grid that lives in the global state is read once per tick, not twice */
    float fPhase = (float)t.phase + (angularSpeed * dt);
    if (fPhase > 100000.0) fPhase = (float)(fPhase - 100000.0); else if (fPhase < 0.0) fPhase = (float)(fPhase + 100000.0);
    t.phase = fPhase;
    const float fCoreTempInput = tmaxf(ambient, coreTInput);
    float coreTemp = t.coreTemp;
    coreTemp += ((fCoreTempInput - coreTemp) * (P.internalCoreTransfer * dt));
    const float fAmbientFactor = ((((carSpeed * carSpeed) * P.coolFactorGain) + 1.0f) * P.surfaceTransfer) * dt;
    const float fPctDt = P.patchCoreTransfer * dt;
    const float kSurf = P.surfaceTransfer * dt;
    const float kPatch = P.patchTransfer * dt;
    /* element loop kept rolled (code size); neighbours are visited in the reference's construction order:
     * j == 0: up, down, right, wrap(last) ; j > 0: up, left, down, right-or-wrap(first) */
    PD_UNROLL
    for (int i = 0; i < PD_THERMAL_STRIPES; ++i) {
        const float inj = inBase + (i == 0 ? in0 : (i == 1 ? in1 : in2));
        const int pi = i * PD_THERMAL_ELEMENTS;
        float sPre = 0;
        PD_NOUNROLL
        for (int j = 0; j < PD_THERMAL_ELEMENTS; ++j) {
            const float fInputT = (j == inElem) ? inj : inBase;
            float fPatchT = g.get(pi + j);
            sPre += fPatchT;
            if (fInputT <= ambient) fPatchT += ((ambient - fPatchT) * fAmbientFactor);
            else fPatchT += ((fInputT - fPatchT) * kSurf);
            if (i > 0) fPatchT += (g.get(pi + j - PD_THERMAL_ELEMENTS) - fPatchT) * kPatch;
            if (j == 0) {
                if (i + 1 < PD_THERMAL_STRIPES) fPatchT += (g.get(pi + PD_THERMAL_ELEMENTS) - fPatchT) * kPatch;
                fPatchT += (g.get(pi + 1) - fPatchT) * kPatch;
                fPatchT += (g.get(pi + PD_THERMAL_ELEMENTS - 1) - fPatchT) * kPatch;
            } else {
                fPatchT += (g.get(pi + j - 1) - fPatchT) * kPatch;
                if (i + 1 < PD_THERMAL_STRIPES) fPatchT += (g.get(pi + j + PD_THERMAL_ELEMENTS) - fPatchT) * kPatch;
                fPatchT += (g.get(pi + ((j + 1 < PD_THERMAL_ELEMENTS) ? j + 1 : 0)) - fPatchT) * kPatch;
            }
            fPatchT += (coreTemp - fPatchT) * fPctDt;
            g.set(pi + j, fPatchT);
            coreTemp += ((fPatchT - coreTemp) * fPctDt);
        }
        if (preSum) preSum[i] = sPre;
    }
    t.coreTemp = coreTemp;
    if (P.performanceCurve.n > 0) {
        /* getCurrentCPTemp (TyreThermalModel.cpp:168-180) with the NEW phase */
        const float fNormCsk = tclampf((camberRAD * P.camberSpreadK), -1.0f, 1.0f);
        const float ph = (float)(t.phase * 0.1591549430964443);
        const int iElemY = ((int)(ph * PD_THERMAL_ELEMENTS)) % PD_THERMAL_ELEMENTS;
        float t0 = 0, t1 = 0, t2 = 0;
        if (iElemY >= 0 && iElemY < PD_THERMAL_ELEMENTS) { t0 = g.get(iElemY); t1 = g.get(iElemY + PD_THERMAL_ELEMENTS); t2 = g.get(iElemY + 2 * PD_THERMAL_ELEMENTS); }
        const float cp = ((((fNormCsk + 1.0f) * t0) + t1) + ((1.0f - fNormCsk) * t2)) * 0.33333334f;
        const float fPracT = ((cp - coreTemp) * 0.25f) + coreTemp;
        t.practicalTemp = fPracT;
        t.thermalMultD = curve_value(P.performanceCurve, fPracT);
    }
}

/* Tyre::step (Tyre.cpp:427-659) with addGroundContact (:661-723), addTyreForcesV10 (TyreForces.cpp:13-194),
 * updateLockedState/updateAngularSpeed (Tyre.cpp:725-752), stepThermalModel (:766-815), stepGrainBlister
 * (:829-922, consumption rate 0 branch), stepFlatSpot (:924-950).
 * `hubBody` is the body the wheel's forces go to (strut hub or the rigid axle). */
template <class SVX> PD_HDN void tyre_step(const PdCarParams& PP, const TrackDev& T, int w, const CarCtx& X, const SVX& sv, Body& hubBody, const Frame& hubFrame, Body& C, float brakeTorqueIn, float handBrakeIn, WheelLink& L) {
    const PdTyre& P = PP.tyre[w];
    const float dt = X.dt;
    TyreS tLocal; TyreS* tp = &tLocal;
    if constexpr (sv_traits<SVX>::in_place) tp = tyre_in_place(sv, w); else load_tyre(sv, w, tLocal);
    TyreS& t = *tp;
    if constexpr (sv_traits<SVX>::grid_in_place) grid_prefetch(sv, PD_OFF_TYRE_PATCH(w));
    if constexpr (sv_traits<SVX>::grid_in_place) { if (w + 1 < 4) tyre_prefetch_l2(sv, w + 1); }
    t.brakeTorque = brakeTorqueIn; t.handBrakeTorque = handBrakeIn;
    t.feedbackTorque = 0; t.Fx = 0; t.Mz = 0; t.slipFactor = 0; t.rollingResistence = 0;
    t.slidingVelocityY = 0; t.slidingVelocityX = 0; t.totalHubVelocity = 0;
    t.surfaceId = -1; t.hasContact = 0;
    const V3 vWorldM2 = hubFrame.ay;
    const V3 worldPosition = hubFrame.p;
    if (!finitef(t.angularVelocity)) t.angularVelocity = 0;

    const RayHit hit = ray_cast_down(T, v3(worldPosition.x, worldPosition.y + 2.0f, worldPosition.z), 3.0f);
    PD_PHASE(X, 3);
    float gripMod = 0, dirtAdditiveK = 0;
    bool punctureCheck = false;
    bool contact = hit.hit && !(hubFrame.ay.y <= 0.35f);
    if (!contact) {
        t.ndSlip = 0; t.Fy = 0;
    } else {
        const PdSurface surf = T.surfaces[hit.surface];
        gripMod = surf.gripMod; dirtAdditiveK = surf.dirtAdditiveK;
        t.surfaceId = hit.surface; t.hasContact = 1;
        V3 vHitPos = hit.pos, vHitNorm = hit.normal;
        const float fTest = dot(vHitNorm, vWorldM2);
        if (fTest <= 0.96f) {
            float fTestAcos;
            if (fTest <= -1.0f || fTest >= 1.0f) fTestAcos = 0; else fTestAcos = m_acos(fTest);
            const float fAngle = fTestAcos - m_acos(0.96f);
            const V3 vAxis = v3((vWorldM2.z * vHitNorm.y) - (vWorldM2.y * vHitNorm.z), (vWorldM2.x * vHitNorm.z) - (vWorldM2.z * vHitNorm.x), (vWorldM2.y * vHitNorm.x) - (vWorldM2.x * vHitNorm.y));
            const M33 m = axis_angle(norm(vAxis), fAngle);
            vHitNorm = v3((((m.m11 * vHitNorm.x) + (m.m21 * vHitNorm.y)) + (m.m31 * vHitNorm.z)) + 0.0f,
                          (((m.m12 * vHitNorm.x) + (m.m22 * vHitNorm.y)) + (m.m32 * vHitNorm.z)) + 0.0f,
                          (((m.m13 * vHitNorm.x) + (m.m23 * vHitNorm.y)) + (m.m33 * vHitNorm.z)) + 0.0f);
        } else {
            const V3 vHitOff = vHitPos - worldPosition;
            const float fDot = dot(vHitNorm, vHitOff);
            vHitPos = (vHitNorm * fDot) + worldPosition;
        }
        V3 contactPoint = vHitPos; const V3 contactNormal = vHitNorm;
        if (surf.sinHeight != 0.0f) {
            const float L = surf.sinLength;
            contactPoint.y -= (((m_sin(L * contactPoint.x) * m_cos(L * contactPoint.z)) + 1.0f) * surf.sinHeight);
        }
        if (surf.granularity != 0.0f) {
            const float v1[3] = {1.0f, 5.8f, 11.4f}; const float v2[3] = {0.005f, 0.005f, 0.01f};
            const float cx = contactPoint.x, cz = contactPoint.z; float cy = contactPoint.y;
            for (int id = 0; id < 3; ++id) { const float v = v1[id]; cy = cy + ((((m_sin(v * cx) * m_cos(v * cz)) + 1.0f) * v2[id]) * -0.6f); }
            contactPoint.y = cy;
        }
        t.contactX = contactPoint.x; t.contactY = contactPoint.y; t.contactZ = contactPoint.z;
        t.normalX = contactNormal.x; t.normalY = contactNormal.y; t.normalZ = contactNormal.z;

        /* ---- addGroundContact ---- */
        {
            const V3 vOffset = worldPosition - contactPoint;
            const float fDistToGround = len(vOffset);
            t.distToGround = fDistToGround;
            float fRadius;
            if (P.radiusRaiseK == 0.0f) fRadius = P.radius; else fRadius = (fabsf(t.angularVelocity) * P.radiusRaiseK) + P.radius;
            if (t.inflation < 1.0f) fRadius = ((fRadius - P.rimRadius) * t.inflation) + P.rimRadius;
            t.liveRadius = fRadius; t.effectiveRadius = fRadius;
            if (fDistToGround > fRadius) {
                t.loadedRadius = fRadius; t.depth = 0; t.load = 0; t.Fy = 0; t.Fx = 0; t.Mz = 0; t.ndSlip = 0;
            } else {
                const float fDepth = fRadius - fDistToGround;
                const float fLoadedRadius = fRadius - fDepth;
                t.depth = fDepth; t.loadedRadius = fLoadedRadius;
                float fMaybePressure;
                if (fLoadedRadius <= P.rimRadius) fMaybePressure = 200000.0f;
                else { fMaybePressure = ((t.pressureDynamic - P.pressureRef) * P.pressureSpringGain) + P.k; if (fMaybePressure < 0.0f) fMaybePressure = 0; }
                const V3 vHubVel = body_point_vel(hubBody, contactPoint);
                const float fLoad = -((dot(vHubVel, contactNormal)) * P.d) + (fDepth * fMaybePressure);
                t.load = fLoad;
                add_force_at_pos(hubBody, contactNormal * fLoad, contactPoint);
                if (t.load < 0.0f) t.load = 0;
            }
        }
        /* ---- addTyreForcesV10 ---- */
        {
            const V3 pos = contactPoint, normal = contactNormal;
            V3 vNegM3 = hubFrame.az * -1.0f;
            V3 roadHeading = norm(vNegM3 - normal * dot(vNegM3, normal));
            const V3 vM1 = hubFrame.ax;
            V3 roadRight = norm(vM1 - normal * dot(vM1, normal));
            const V3 hubAngVel = hubBody.w;
            const V3 hubPointVel = body_point_vel(hubBody, pos);
            t.slidingVelocityY = dot(hubPointVel, roadRight);
            const float roadVelocityX = -(dot(hubPointVel, roadHeading));
            float fSlipAngleTmp = (roadVelocityX != 0.0f) ? atanf(-(t.slidingVelocityY / fabsf(roadVelocityX))) : 0.0f;
            const float fTmp = dot(hubAngVel, vM1) + t.angularVelocity;
            t.slidingVelocityX = (fTmp * t.effectiveRadius) - roadVelocityX;
            const float fRoadVelocityXAbs = fabsf(roadVelocityX);
            float fSlipRatioTmp = ((fRoadVelocityXAbs == 0.0f) ? 0.0f : (t.slidingVelocityX / fRoadVelocityXAbs));
            { /* calcCamberRAD (TyreUtils.inl:14-21) */
                const float f = ((hubFrame.ax.y * contactNormal.y) + (hubFrame.ax.x * contactNormal.x)) + (hubFrame.ax.z * contactNormal.z);
                t.camberRAD = (f <= -1.0f || f >= 1.0f) ? -1.5707964f : -m_asin(f);
            }
            t.totalHubVelocity = sqrtf((roadVelocityX * roadVelocityX) + (t.slidingVelocityY * t.slidingVelocityY));
            const float fNdSlip = tclampf(t.ndSlip, 0.0f, 1.0f);
            const float fLoadDivFz0 = t.load / P.Fz0;
            const float fRelaxLen = P.relaxationLength;
            const float fRelax1 = (((fLoadDivFz0 * fRelaxLen) - fRelaxLen) * 0.3f) + fRelaxLen;
            const float fRelax2 = ((fRelaxLen - (fRelax1 * 2.0f)) * fNdSlip) + (fRelax1 * 2.0f);
            if (t.totalHubVelocity < 1.0f) {
                fSlipRatioTmp = tclampf(t.slidingVelocityX * 0.5f, -1.0f, 1.0f);
                fSlipAngleTmp = tclampf(t.slidingVelocityY * -5.5f, -1.0f, 1.0f);
            }
            const float fSlipRatio = t.slipRatio;
            const float fSlipRatioDelta = fSlipRatioTmp - fSlipRatio;
            float fNewSlipRatio = fSlipRatioTmp;
            if (fRelax2 != 0.0f) {
                const float sc = (t.totalHubVelocity * dt) / fRelax2;
                if (sc <= 1.0f) { if (sc < 0.04f) fNewSlipRatio = (0.04f * fSlipRatioDelta) + fSlipRatio; else fNewSlipRatio = (sc * fSlipRatioDelta) + fSlipRatio; }
            }
            const float fSlipAngle = t.slipAngleRAD;
            const float fSlipAngleDelta = fSlipAngleTmp - fSlipAngle;
            float fNewSlipAngle = fSlipAngleTmp;
            if (fRelax2 != 0.0f) {
                const float sc = (t.totalHubVelocity * dt) / fRelax2;
                if (sc <= 1.0f) { if (sc < 0.04f) fNewSlipAngle = (0.04f * fSlipAngleDelta) + fSlipAngle; else fNewSlipAngle = (sc * fSlipAngleDelta) + fSlipAngle; }
            }
            t.slipAngleRAD = fNewSlipAngle; t.slipRatio = fNewSlipRatio;
            if (t.load <= 0.0f) { t.slipAngleRAD = 0; t.slipRatio = 0; }
            /* getCorrectedD(1.0, &wearMult) (TyreForces.cpp:196-210) */
            float fCorrectedD = (1.0f * t.thermalMultD) / ((fabsf(t.pressureDynamic - P.idealPressure) * P.pressureGainD) + 1.0f);
            if (P.wearCurve.n) { const float wm = curve_value(P.wearCurve, (float)t.virtualKM); fCorrectedD *= wm; t.wearMult = wm; }
            TmIn tmi;
            tmi.load = t.load; tmi.slipAngleRAD = t.slipAngleRAD; tmi.slipRatio = t.slipRatio; tmi.camberRAD = t.camberRAD;
            tmi.speed = t.totalHubVelocity; tmi.u = (fCorrectedD * gripMod) * T.info.dynamicGripLevel;
            { /* calcContactPatchLength (TyreUtils.inl:23-30) */
                const float v = t.liveRadius - t.depth;
                tmi.cpLength = (v <= 0.0f || t.liveRadius <= v) ? 0.0f : sqrtf((t.liveRadius * t.liveRadius) - (v * v)) * 2.0f;
            }
            tmi.grain = 0.0f; tmi.blister = 0.0f;
            tmi.pressureRatio = (t.pressureDynamic / P.idealPressure) - 1.0f;
            const TmOut tmo = sctm_solve(P, tmi);
            t.Fy = tmo.Fy * 1.0f /* aiMult */; t.Fx = -tmo.Fx; t.Dy = tmo.Dy; t.Dx = tmo.Dx;
            float fHubSpeed = t.effectiveRadius * t.angularVelocity;
            { /* stepDirtyLevel (TyreForces.cpp:212-233) */
                const float hs = fabsf(fHubSpeed);
                if (t.dirtyLevel < 5.0f) t.dirtyLevel += (((hs * dirtAdditiveK) * 0.03f) * dt);
                if (dirtAdditiveK == 0.0f) { if (t.dirtyLevel > 0.0f) t.dirtyLevel -= ((hs * 0.015f) * dt); if (t.dirtyLevel < 0.0f) t.dirtyLevel = 0; }
                const float fM = tmaxf(0.8f, (1.0f - tclampf(t.dirtyLevel * 0.05f, 0.0f, 1.0f)));
                t.Fy *= fM; t.Fx *= fM; t.Mz *= fM;
            }
            if (PP.mechanicalDamageRate > 0.0f) { /* stepPuncture (TyreForces.cpp:235-247) */
                if constexpr (sv_traits<SVX>::grid_in_place && PD_PUNCTURE_FUSED) punctureCheck = true;       /* deferred to the thermal sweep below, which reads the same 36 words (inflation is not read again this tick) */
                else {
                    bool expl = false;
                    PD_UNROLL
                    for (int i = 0; i < 3; ++i) { float s = 0; PD_UNROLL for (int j = 0; j < 12; ++j) { if constexpr (sv_traits<SVX>::grid_in_place) s += sv.f(PD_OFF_TYRE_PATCH(w) + j + i * 12); else s += t.T[j + i * 12]; } if (s / 12.0f > P.explosionTemperature) expl = true; }
                    if (expl) t.inflation = 0;
                }
            }
            t.Mz = tmo.Mz;
            V3 vForce = (roadHeading * t.Fx) + (roadRight * t.Fy);
            if (!(finitef(vForce.x) && finitef(vForce.y) && finitef(vForce.z))) vForce = v3(0, 0, 0);
            add_force_at_pos(hubBody, vForce, pos);
            t.localMX = -(t.loadedRadius * t.Fx);
            add_torque(hubBody, normal * tmo.Mz);
            const float fAngularVelocityAbs = fabsf(t.angularVelocity);
            if (fAngularVelocityAbs > 1.0f) {
                fHubSpeed = t.effectiveRadius * t.angularVelocity;
                const float fHubSpeedSign = signf_(fHubSpeed);
                const float fPressureDynamic = t.pressureDynamic;
                float fPressureUnk = (((P.idealPressure / fPressureDynamic) - 1.0f) * P.pressureRRGain) + 1.0f;
                if (fPressureDynamic <= 0.0f) fPressureUnk = 0;
                float fRrUnk = ((((fHubSpeed * fHubSpeed) * P.rr1) + P.rr0) * fHubSpeedSign) * fPressureUnk;
                if (fAngularVelocityAbs > 20.0f) {
                    const float fNdSlipNorm = tclampf(t.ndSlip, 0.0f, 1.0f);
                    const float fRrSlipUnk = fPressureUnk * P.rr_slip;
                    const float fSlipUnk = fNdSlipNorm * fRrSlipUnk;
                    fRrUnk = fRrUnk * ((fSlipUnk * 0.001f) + 1.0f);
                }
                t.rollingResistence = -(((t.load * 0.001f) * fRrUnk) * t.effectiveRadius);
            }
            {
                const float fSlidingVelocity = sqrtf(t.slidingVelocityX * t.slidingVelocityX + t.slidingVelocityY * t.slidingVelocityY);
                float fLoadVKM = 1.0f;
                if (P.useLoadForVKM) fLoadVKM = t.load / P.Fz0;
                t.virtualKM += (((fSlidingVelocity * dt) * PP.tyreConsumptionRate) * fLoadVKM) * 0.001f;
            }
            const float fStaticDy = sctm_static_dy(P, t.load);
            t.ndSlip = tmo.ndSlip; t.D = fStaticDy;
        }
        if (surf.damping > 0.0f) { /* Tyre.cpp:591-602 */
            const V3 vForce = C.v * -(C.mass * surf.damping);
            add_force_at_rel_pos(C, vForce, v3(0, 0, 0));
        }
    }
    /* LB_COMPUTE_TORQ */
    const float fHandBrakeTorque = t.handBrakeTorque;
    float fBrakeTorque = t.brakeTorque * 1.0f /* absOverride */;
    if (fBrakeTorque <= fHandBrakeTorque) fBrakeTorque = fHandBrakeTorque;
    const float fAngularVelocitySign = signf_(t.angularVelocity);
    float fTorq = t.rollingResistence - ((fAngularVelocitySign * fBrakeTorque) + t.localMX);
    if (!finitef(fTorq)) fTorq = 0;
    float fFeedbackTorque = fTorq + 0.0f /* electricTorque */;
    if (!finitef(fFeedbackTorque)) fFeedbackTorque = 0;
    t.feedbackTorque = fFeedbackTorque;
    if (P.driven) {
        /* updateLockedState with driven == true always clears isLocked when it was set (Tyre.cpp:725-736) */
        if (t.isLocked) {
            const float fBrake = tmaxf(1.0f * t.brakeTorque, t.handBrakeTorque);
            t.isLocked = (fabsf(fBrake) >= fabsf(t.loadedRadius * t.Fx)) && (fabsf(t.angularVelocity) < 1.0f) && (!P.driven);
        }
        const float fS0 = signf_(t.oldAngularVelocity), fS1 = signf_(t.angularVelocity);
        if (fS0 != fS1 && t.totalHubVelocity < 1.0f) t.isLocked = 1;
        t.oldAngularVelocity = t.angularVelocity;
    } else {
        /* updateAngularSpeed (Tyre.cpp:738-752) */
        if (t.isLocked) {
            const float fBrake = tmaxf(1.0f * t.brakeTorque, t.handBrakeTorque);
            t.isLocked = (fabsf(fBrake) >= fabsf(t.loadedRadius * t.Fx)) && (fabsf(t.angularVelocity) < 1.0f) && (!P.driven);
        }
        const float fAngVel = t.angularVelocity + ((t.feedbackTorque / P.angularInertia) * dt);
        if (signf_(fAngVel) != signf_(t.oldAngularVelocity)) t.isLocked = 1;
        t.oldAngularVelocity = fAngVel;
        t.angularVelocity = t.isLocked ? 0.0f : fAngVel;
        if (fabsf(t.angularVelocity) < 1.0f) t.angularVelocity *= 0.9f;
        /* stepRotationMatrix only spins the visual wheel matrix: not state of the path */
    }
    if (t.totalHubVelocity < 10.0f) t.slipFactor = fabsf(t.totalHubVelocity * 0.1f) * t.slipFactor;

    PD_PHASE(X, 4);
    /* ---- stepThermalModel (Tyre.cpp:766-815) ---- */
    {
        float fThermalInput = (sqrtf((t.slidingVelocityX * t.slidingVelocityX) + (t.slidingVelocityY * t.slidingVelocityY)) * ((t.D * t.load) * P.thermalFrictionK)) * T.info.dynamicGripLevel;
        if (t.surfaceId >= 0) fThermalInput *= gripMod;
        t.thermalInput = fThermalInput;
        if (finitef(fThermalInput)) {
            const float fPressureDynamic = t.pressureDynamic, fIdealPressure = P.idealPressure;
            float fThermalRollingK = P.thermalRollingK;
            const float fScale = (((fIdealPressure / fPressureDynamic) - 1.0f) * P.pressureRRGain) + 1.0f;
            if (fPressureDynamic >= 0.0) fThermalRollingK *= fScale;
            if (P.version < 5) t.thermalInput += (((fThermalRollingK * t.angularVelocity) * t.load) * 0.001f);
            if (P.version >= 6) t.thermalInput += ((((fScale * P.thermalRollingSurfaceK) * t.angularVelocity) * t.load) * 0.001f);
            const float inBase = X.c.thermalPrimed ? 0.0f : PP.ambientTemperature;
            int inElem; float in0, in1, in2;
            { /* addThermalInput (TyreThermalModel.cpp:147-166), phase BEFORE this tick's update */
                const float xpos = t.camberRAD, pressureRel = (fPressureDynamic / fIdealPressure) - 1.0f, temp = t.thermalInput;
                const float fNormXcs = tclampf((xpos * P.camberSpreadK), -1.0f, 1.0f);
                const float ph = (float)(t.phase * 0.1591549430964443);
                inElem = ((int)(ph * PD_THERMAL_ELEMENTS)) % PD_THERMAL_ELEMENTS;
                const float fT = PP.roadTemperature + temp;
                const float fPr1 = pressureRel * 0.1f, fPr2 = (pressureRel * -0.5f) + 1.0f;
                in0 = ((((fNormXcs + 1.0f) - (fPr1 * 0.5f)) * fPr2) * fT);
                in1 = (((fPr1 + 1.0f) * fPr2) * fT);
                in2 = ((((1.0f - fNormXcs) - (fPr1 * 0.5f)) * fPr2) * fT);
            }
            float coreTInput = 0.0f;
            if (P.version >= 5) coreTInput += (((fThermalRollingK * t.angularVelocity) * t.load) * 0.001f);
            if constexpr (sv_traits<SVX>::grid_in_place) {
                float ss[3];
                thermal_step(P, t, GridSV<SVX>{sv, PD_OFF_TYRE_PATCH(w)}, inBase, inElem, in0, in1, in2, coreTInput, dt, t.angularVelocity, t.camberRAD, PP.ambientTemperature, X.c.speed, ss);
                if (punctureCheck) { punctureCheck = false; if (ss[0] / 12.0f > P.explosionTemperature || ss[1] / 12.0f > P.explosionTemperature || ss[2] / 12.0f > P.explosionTemperature) t.inflation = 0; }
            } else thermal_step(P, t, GridLocal{t.T}, inBase, inElem, in0, in1, in2, coreTInput, dt, t.angularVelocity, t.camberRAD, PP.ambientTemperature, X.c.speed);
        }
        if constexpr (sv_traits<SVX>::grid_in_place) {
            if (punctureCheck) {        /* no thermal sweep this tick (non-finite thermal input): stepPuncture (TyreForces.cpp:235-247) on its own */
                bool expl = false;
                PD_UNROLL
                for (int i = 0; i < 3; ++i) { float s = 0; PD_UNROLL for (int j = 0; j < 12; ++j) s += sv.f(PD_OFF_TYRE_PATCH(w) + j + i * 12); if (s / 12.0f > P.explosionTemperature) expl = true; }
                if (expl) t.inflation = 0;
            }
        }
    }
    t.pressureDynamic = ((t.coreTemp - 26.0f) * P.pressureTemperatureGain) + t.pressureStatic;
    /* stepGrainBlister: tyreConsumptionRate == 0 -> grain = blister = 0 (Tyre.cpp:917-921); rate > 0 is rejected at create */
    /* stepFlatSpot (Tyre.cpp:924-950) */
    if (fabsf(t.angularVelocity) <= 0.3f || t.slipRatio < -0.98f) {
        if (t.surfaceId >= 0 && t.totalHubVelocity > 3.0f) {
            const float fDamage = PP.mechanicalDamageRate;
            if (fDamage != 0.0f && gripMod >= 0.95f) {
                t.flatSpot += (((t.totalHubVelocity * P.flatSpotK) * t.load) * gripMod) * 0.00001f * dt * fDamage * P.softnessIndex;
                if (t.flatSpot > 1.0f) t.flatSpot = 1.0f;
            }
        }
    }
    if constexpr (!sv_traits<SVX>::in_place) store_tyre(sv, w, t);
    L.load = t.load; L.feedbackTorque = t.feedbackTorque; L.angularVelocity = t.angularVelocity;
    L.brakeTorque = t.brakeTorque; L.handBrakeTorque = t.handBrakeTorque; L.ndSlip = t.ndSlip; L.slipRatio = t.slipRatio;
    L.isLocked = t.isLocked; L.surfaceId = t.surfaceId;
}

/* ================================ aero ================================ */
/* AeroMap::step -> Wing::step / addDrag / addLift (AeroMap.cpp:83-96, Wing.cpp:71-204); wind is zero
 * (Simulator::stepWind is compiled out, Simulator.cpp:203-224) so getGroundWindVector contributes 0 */
/* wings first, first + step, ...: the quad kernel gives each lane its own wings (their forces meet in the quad-wide
 * sum of the chassis force), the thread-per-car kernel passes (0, 1) */
PD_HD void aero_step(const PdCarParams& PP, Body& C, int first = 0, int step = 1) {
    for (int i = first; i < PP.nWings; i += step) {
        const PdWing& W = PP.wing[i];
        const V3 pos = v3(W.position[0], W.position[1], W.position[2]);
        const V3 vWorldVel = body_rel_point_vel(C, pos);
        const V3 lv = irot(C.fr, vWorldVel + v3(0, 0, 0));
        if (lv.z == 0.0f) continue;
        const float aoa = atanf((1.0f / lv.z) * lv.y) * 57.29578f;
        const float yawAngle = atanf((1.0f / lv.z) * lv.x) * 57.29578f;
        { /* addDrag */
            const float fAngleOff = W.isVertical ? yawAngle : aoa;
            const float cd = curve_value(W.lutAOA_CD, (W.angleMult * W.angle) + fAngleOff) * W.cdGain;
            const float fDot = sqlen(lv);
            const float fDrag = (((fDot * cd) * PP.airDensity) * W.area) * 0.5f;
            if (fDot != 0.0f) add_rel_force_at_rel_pos(C, norm(lv) * -fDrag, pos);
        }
        { /* addLift */
            float fAngleOff, fAxis;
            if (W.isVertical) { fAngleOff = yawAngle; fAxis = lv.x; } else { fAngleOff = aoa; fAxis = lv.y; }
            float cl = curve_value(W.lutAOA_CL, (W.angleMult * W.angle) + fAngleOff) * W.clGain;
            if (lv.z < 0.0f) cl = 0;
            if (!W.isVertical && W.yawGain != 0.0f) { const float v8 = (m_sin(fabsf(yawAngle) * 0.017453f) * W.yawGain) + 1.0f; cl *= tclampf(v8, 0.0f, 1.0f); }
            const float fDot = (fAxis * fAxis) + (lv.z * lv.z);
            const float fLift = (((fDot * cl) * PP.airDensity) * W.area) * 0.5f;
            if (fDot != 0.0f) {
                const V3 vNorm = norm(lv);
                const V3 vOut = W.isVertical ? v3(-vNorm.z, 0, vNorm.x) : v3(0, vNorm.z, -vNorm.y);
                add_rel_force_at_rel_pos(C, vOut * -fLift, pos);
            }
        }
    }
}

/* ================================ assists / gearbox / engine / drivetrain ================================ */

/* event fired by Drivetrain::gearUp/gearDown -> AutoClutch::onGearRequest (AutoClutch.cpp:202-221) and the
 * AutoBlip lambda (AutoBlip.cpp:38-48).  request: 1 = up, 2 = down */
PD_HD void on_gear_request(const PdCarParams& PP, CarCtx& X, int request) {
    CarS& c = X.c; const PdAssists& A = PP.assists;
    if (A.acUseAutoOnChange) {
        if (request == 2 && c.acClutchValueSignal > 0.01f && A.downshiftProfile.n == 4) { c.acSeqProfile = 2; c.acSeqTime = 0.0f; c.acSeqDone = 0; }
        if (request == 1 && c.acClutchValueSignal > 0.01f && A.upshiftProfile.n == 4) { c.acSeqProfile = 1; c.acSeqTime = 0.0f; c.acSeqDone = 0; }
    }
    if (request == 2) { if (c.ctlClutch > 0.1f) c.blipStartTime = (X.time * 1000.0); }
}

/* AutoClutch::step / stepSequence (AutoClutch.cpp:91-200) */
PD_HD void autoclutch_step(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c; const PdAssists& A = PP.assists; const float dt = X.dt;
    if (!c.acSeqDone) {
        const bool bIsMoving = (c.speed * 3.6f) >= 5.0f;
        if (bIsMoving) {
            const PdCurve& cur = (c.acSeqProfile == 2) ? A.downshiftProfile : A.upshiftProfile;
            const float sig = (c.acSeqProfile == 0) ? 0.0f : curve_value(cur, c.acSeqTime);
            c.acClutchValueSignal = sig;
            c.acSeqTime += dt;
            const float maxRef = (c.acSeqProfile == 0 || cur.n <= 0) ? 0.0f : cur.ref[cur.n - 1];
            if (c.acSeqTime > maxRef) c.acSeqDone = 1;
            c.ctlClutch = tclampf(c.acClutchValueSignal, 0.0f, 1.0f);
            return;
        }
        c.acSeqDone = 1;
    }
    if (A.acUseAutoOnStart || A.acIsForced) {
        float fNewClutchInput = 1.0f, fNewSignal = 1.0f;
        const float fEngineRpm = engine_rpm(c);
        const int iCurGear = c.currentGear;
        const bool bIsStationary = (c.speed * 3.6f) < 5.0f;
        bool toL20 = false;
        if ((iCurGear & 0xFFFFFFFD) != 0) {
            if (iCurGear == 1) {
                if (!bIsStationary) { /* LABEL_21 */ }
                else if (c.ctlGas > 0.2f) { c.acClutchValueSignal = 1.0f; }
                else toL20 = true;
            } else {
                toL20 = fEngineRpm < A.acRpmMin;
            }
        } else {
            if (fEngineRpm >= A.acRpmMin) {
                if (fEngineRpm <= A.acRpmMax) { fNewSignal = (fEngineRpm - A.acRpmMin) / (A.acRpmMax - A.acRpmMin); c.acClutchValueSignal = fNewSignal; }
            }
            if (fEngineRpm > A.acRpmMax) fNewSignal = 1.0f;
            toL20 = fEngineRpm < A.acRpmMin;
        }
        if (toL20) { fNewSignal = 0.0f; c.acClutchValueSignal = 0.0f; }
        /* LABEL_21 */
        const float fSig = c.acClutchValueSignal;
        const float fStep = dt * A.acClutchSpeed;
        if (fabsf(fNewSignal - fSig) >= fStep) { if (fNewSignal <= fSig) c.acClutchValueSignal = fSig - fStep; else c.acClutchValueSignal = fStep + fSig; }
        else c.acClutchValueSignal = fNewSignal;
        if (c.acClutchValueSignal <= 1.0f) { if (c.acClutchValueSignal >= 0.0f) fNewClutchInput = c.acClutchValueSignal; else fNewClutchInput = 0.0f; }
        c.ctlClutch = fNewClutchInput;
    }
}

/* AutoBlip::step (AutoBlip.cpp:51-75) */
PD_HD void autoblip_step(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c; const PdAssists& A = PP.assists;
    if ((A.blipIsActive || A.blipIsElectronic) && ((c.speed * 3.6f) > 5.0f)) {
        const double fBlipElapsed = (X.time * 1000.0) - c.blipStartTime;
        if (fBlipElapsed >= 0.0 && fBlipElapsed < A.blipPerformTime && A.blipProfile.n == 4) {
            float fGas = c.ctlGas;
            const float fProfile = curve_value(A.blipProfile, (float)fBlipElapsed);
            if (fGas <= fProfile) fGas = fProfile;
            float fNewGas = 1.0;
            if (fGas <= 1.0) { fNewGas = 0.0; if (fGas >= 0.0) fNewGas = fGas; }
            c.ctlGas = fNewGas;
        }
    }
}

/* AutoShifter::step (AutoShifter.cpp:31-117); changeUpRpm / changeDnRpm come resolved from the loader */
PD_HD void autoshift_step(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c; const PdAssists& A = PP.assists;
    if (!A.asIsActive) return;
    if (c.currentGear) {
        if (!c.ctlGearUp && !c.ctlGearDn) {
            c.ctlGearDn = 0; c.ctlGearUp = 0;
            bool bIsSlipping = false;
            const float drivingSlip = (PP.drivetrain.tractionType == 1) ? tmaxf(X.wl[0].ndSlip, X.wl[1].ndSlip) : tmaxf(X.wl[2].ndSlip, X.wl[3].ndSlip);
            if (drivingSlip > A.asSlipThreshold) { if (c.speed > 5.0f) bIsSlipping = true; }
            const bool changing = c.reqRequest != 0;
            if (!changing) {
                if ((c.ctlClutch > 0.99f || c.currentGear == 1) && !bIsSlipping) {
                    const int iEngineRpm = (int)engine_rpm(c);
                    if (iEngineRpm > A.asChangeUpRpm) {
                        if (c.currentGear < (PP.drivetrain.nGears - 1) && c.ctlGas > 0.2f && c.gasCutoff <= 0.0f) { c.ctlGearUp = 1; c.gasCutoff = A.asGasCutoffTime; }
                    }
                    const int iCurGear = c.currentGear;
                    int iChangeDnRpm;
                    if (iCurGear == 3) iChangeDnRpm = (int)(A.asChangeDnRpm * 0.65f); else iChangeDnRpm = A.asChangeDnRpm;
                    if (iEngineRpm < iChangeDnRpm && iCurGear > 2 && c.ctlClutch > 0.85f && c.gasCutoff <= 0.0f) c.ctlGearDn = 1;
                }
            }
            const bool bLowSpeed = c.speed < 2.0f;
            if (bLowSpeed && !changing) { if (c.ctlGas < 0.1f && c.gasCutoff <= 0.0f && c.currentGear > 2) c.ctlGearDn = 1; }
            const float fCutoff = c.gasCutoff;
            if (fCutoff > 0.0f) { c.gasCutoff = fCutoff - X.dt; c.ctlGas = 0.0f; }
        }
    }
}

/* Drivetrain::setCurrentGear (Drivetrain.cpp:174-210) */
PD_HD void dt_set_current_gear(const PdCarParams& PP, CarCtx& X, int index, bool force) {
    CarS& c = X.c; const PdDrivetrain& D = PP.drivetrain;
    c.isGearGrinding = 0;
    if (index >= 0 && index < D.nGears && index != c.currentGear) {
        const double v8 = fabs(c.engineVel - D.gears[index] * c.driveVel * D.finalRatio);
        const double v9 = ((c.ctlGas * c.locClutch * v8 - c.locClutch * v8) * D.controlsWindowGain + c.locClutch * v8) * (1.0f / (2.0f * 3.1415926535897932384626433832795f)) * 60.0f;
        if (index == 1 || force || v9 < c.validShiftRPMWindow) c.currentGear = index;
        else {
            c.isGearGrinding = 1;
            if (c.validShiftRPMWindow > 0.0) { const double fRate = PP.mechanicalDamageRate; if (fRate > 0.0) c.validShiftRPMWindow -= D.damageRpmWindow * fRate * 0.003; }
        }
    }
}
/* Drivetrain::gearUp / gearDown (Drivetrain.cpp:212-279) */
PD_HD bool dt_gear_up(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c; const PdDrivetrain& D = PP.drivetrain;
    const int iReqGear = c.currentGear + 1;
    if (iReqGear >= D.nGears) return false;
    if (c.reqRequest != 0) return false;
    c.reqRequest = 1; c.reqTimeAcc = 0; c.reqTimeout = D.gearUpTime; c.reqGear = iReqGear;
    on_gear_request(PP, X, 1);
    if (D.autoCutOffTime != 0.0f) c.cutOff = D.autoCutOffTime;
    c.currentGear = 1;
    return true;
}
PD_HD bool dt_gear_down(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c; const PdDrivetrain& D = PP.drivetrain;
    const int iCurGear = c.currentGear, iReqGear = iCurGear - 1;
    if (iCurGear <= 0) return false;
    if (c.reqRequest != 0) return false;
    c.reqRequest = 2; c.reqTimeAcc = 0; c.reqTimeout = D.gearDnTime; c.reqGear = iReqGear;
    on_gear_request(PP, X, 2);
    c.currentGear = 1;
    return true;
}
/* GearChanger::step (GearChanger.cpp:18-40) */
PD_HD void gearchanger_step(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c;
    const int gearId = c.ctlRequestedGear;
    if (gearId == -1) {
        if (c.ctlGearUp && !c.lastGearUp) dt_gear_up(PP, X);
        if (c.ctlGearDn && !c.lastGearDn) dt_gear_down(PP, X);
        c.lastGearUp = c.ctlGearUp ? 1 : 0; c.lastGearDn = c.ctlGearDn ? 1 : 0;
    } else dt_set_current_gear(PP, X, gearId, false);
}

/* Engine::step (Engine.cpp:193-342); no turbos, no coast/torque generators, no restrictor on the demo car */
PD_HD void engine_step(const PdCarParams& PP, CarCtx& X, float gasInput, float rpm) {
    CarS& c = X.c; const PdEngine& E = PP.engine;
    float gas;
    if (E.throttleResponseCurve.n && E.throttleResponseCurveMax.n) {       /* Engine::getThrottleResponseGas (Engine.cpp:344-366) */
        const float fTrc = tclampf(curve_value(E.throttleResponseCurve, gasInput * 100.0f) * 0.01f, 0.0f, 1.0f);
        const float fTrcMax = tclampf(curve_value(E.throttleResponseCurveMax, gasInput * 100.0f) * 0.01f, 0.0f, 1.0f);
        const float fTrcScale = tclampf(rpm / E.throttleResponseCurveMaxRef, 0.0f, 1.0f);
        gas = ((fTrcMax - fTrc) * fTrcScale) + fTrc;
    } else if (E.throttleResponseCurve.n) gas = tclampf(curve_value(E.throttleResponseCurve, gasInput * 100.0f) * 0.01f, 0.0f, 1.0f); else gas = gasInput;
    if (E.gasCoastOffset > 0.0f) {
        float g1 = tclampf((rpm - (float)E.minimum) / (float)E.coastEntryRpm, 0.0f, 1.0f);
        gas = tclampf(((1.0f - (E.gasCoastOffset * g1)) * gas) + (E.gasCoastOffset * g1), 0.0f, 1.0f);
    }
    if (E.limiter && (E.limiter * E.limiterMultiplier) < rpm) c.limiterOn = E.limiterCycles;
    if (c.limiterOn > 0) { gas = 0; c.limiterOn--; }
    if (c.lifeLeft <= 0.0f) c.fuelPressure = 0;
    const float fGas = gas * 1.0f /* electronicOverride */;
    c.gasUsage = fGas;
    float fPower = curve_value(E.powerCurve, rpm);
    float fCoastTorq = 0;
    if (E.coast1 != 0.0f) fCoastTorq = (rpm - (float)E.minimum) * E.coast1;
    if (PP.nTurbos > 0) { /* Engine::stepTurbos (Engine.cpp:368-384) + Turbo::step (Turbo.cpp:11-40; the reference passes dt = 0.003) */
        float boost = 0.0f;
        for (int i = 0; i < PP.nTurbos && i < PD_MAX_TURBOS; ++i) {
            const PdTurbo& U = PP.turbo[i];
            float& rotation = (i == 0) ? c.turboRot0 : ((i == 1) ? c.turboRot1 : c.turboRot2);
            float fNewRotation = 0.0f, fLag;
            if (rpm > 0.0f && fGas > 0.0f) fNewRotation = m_pow(tclampf(((fGas * rpm) / U.rpmRef), 0.0f, 1.0f), U.gamma);
            if (fNewRotation <= rotation) fLag = tclampf((0.003f * U.lagDN), 0.0f, 1.0f); else fLag = tclampf((0.003f * U.lagUP), 0.0f, 1.0f);
            rotation += ((fNewRotation - rotation) * fLag);
            if (U.wastegate != 0.0f) { const float fUserWG = U.wastegate * U.userSetting; if ((U.maxBoost * rotation) > fUserWG) rotation = fUserWG / U.maxBoost; }
            boost += ((U.maxBoost * rotation) * c.fuelPressure);
        }
        c.turboBoost = boost;
        if (boost != 0.0f) fPower *= (boost + 1.0f);
    }
    if (E.coast2 != 0.0f) { const float d = rpm - (float)E.minimum; fCoastTorq -= (((d * d) * E.coast2) * signf_(rpm)); }
    fCoastTorq += 0.0f; /* externalCoastTorque */
    if (rpm <= (float)E.minimum) fCoastTorq = 0;
    if (E.turboBoostDamageThreshold != 0.0f && c.turboBoost > E.turboBoostDamageThreshold) c.lifeLeft -= ((((c.turboBoost - E.turboBoostDamageThreshold) * E.turboBoostDamageK) * 0.003f) * PP.mechanicalDamageRate);
    if (E.rpmDamageThreshold != 0.0f && rpm > E.rpmDamageThreshold) c.lifeLeft -= ((((rpm - E.rpmDamageThreshold) * E.rpmDamageK) * 0.003f) * PP.mechanicalDamageRate);
    const float fAirAmount = PP.airDensity * 0.82630974f;
    const float fOutTorq = ((((fPower - fCoastTorq) * fGas) + fCoastTorq) * fAirAmount);
    double outTorque = fOutTorq;
    if (c.fuelPressure > 0.0f) {
        if (rpm >= (float)E.minimum) {
            if (E.overlapGain != 0.0f) {
                const float fOverlap = m_sin((float)X.time * 0.001f * E.overlapFreq * rpm * 0.0003333333333333333f) * 0.5f - 0.5f;
                outTorque = (fOverlap * fabsf(rpm - E.overlapIdealRPM) * E.overlapGain) + fOutTorq;
            }
        } else if (E.isEngineStallEnabled) outTorque = rpm * -0.01f;
        else outTorque = tmaxf(15.0f, fOutTorq);
    }
    if (c.fuelPressure < 1.0f) outTorque = (outTorque - rpm * -0.01f) * c.fuelPressure + rpm * -0.01f;
    c.outTorque = outTorque;
}

PD_HD double dt_inertia_from_wheels(const PdCarParams& PP, const CarS& c, double ratio, double engineInertia) { /* Drivetrain.cpp:662-698 */
    const PdDrivetrain& D = PP.drivetrain;
    const double fRatioSq = ratio * ratio;
    const double fRWD = D.driveInertia + (D.shaftInertiaL + D.shaftInertiaR);
    if (ratio == 0.0) return fRWD;
    if (c.clutchOpenState) return fRWD + (D.clutchInertia * fRatioSq);
    return fRWD + ((D.clutchInertia + engineInertia) * fRatioSq);
}
PD_HD double dt_inertia_from_engine(const PdCarParams& PP, double ratio, double engineInertia) { /* Drivetrain.cpp:700-725 */
    const PdDrivetrain& D = PP.drivetrain;
    if (ratio == 0.0) return engineInertia;
    const double fRWD = D.driveInertia + D.shaftInertiaL + D.shaftInertiaR;
    return fRWD / (ratio * ratio) + D.clutchInertia + engineInertia;
}

/* Drivetrain::step + step2WD (Drivetrain.cpp:281-585) for RWD/FWD with LSD or spool.
 * dl / dr index the driven wheels (2,3 for RWD). */
PD_HDN float drivetrain_step(const PdCarParams& PP, CarCtx& X) {
    CarS& c = X.c; const PdDrivetrain& D = PP.drivetrain; const float dt = X.dt;
    const int dl = (D.tractionType == 1) ? 0 : 2, dr = dl + 1;
    WheelLink& tl = X.wl[dl]; WheelLink& tr = X.wl[dr];
    c.locClutch = m_pow(c.ctlClutch, 1.5f);
    c.currentClutchTorque = 0;
    const int iGearRequest = c.reqRequest - 1;
    if ((!iGearRequest || iGearRequest == 1) && (c.reqTimeout < c.reqTimeAcc)) { c.currentGear = c.reqGear; c.reqRequest = 0; }
    if (c.reqRequest != 0) c.reqTimeAcc += dt;
    const double curGearRatio = D.gears[c.currentGear];
    const double ratio = D.finalRatio * curGearRatio;
    const double engineInertia = PP.engine.inertia;
    if (c.lastRatio != ratio) {
        /* reallignSpeeds (Drivetrain.cpp:610-635) */
        if (ratio != 0.0) {
            const double fDriveVel = c.driveVel;
            if (c.locClutch <= 0.9f) c.rootVel = fDriveVel * ratio;
            else c.rootVel -= (1.0 - engineInertia / dt_inertia_from_engine(PP, ratio, engineInertia)) * (c.rootVel / ratio - fDriveVel) * fabs(ratio);
            const double acc = (c.rootVel / ratio - fDriveVel);
            c.driveVel += acc; c.shaftRVel += acc; c.shaftLVel += acc;
            if (!c.clutchOpenState) c.engineVel = c.rootVel;
        }
        c.lastRatio = ratio;
    }
    float gasInput = 0;
    if (c.cutOff > 0.0) c.cutOff -= dt; else gasInput = c.ctlGas;
    const float rpm = (float)((c.engineVel * 0.15915507) * 60.0);
    engine_step(PP, X, gasInput, rpm);
    const double outTorque = c.outTorque;
    if (c.locClutch < 1.0f) c.clutchOpenState = 1;
    else if (c.engineVel != 0.0) c.clutchOpenState = (fabs(c.rootVel / c.engineVel - 1.0) >= 0.1) ? 1 : 0;
    else c.clutchOpenState = (c.rootVel != 0.0) ? 1 : 0;
    const double fEngineInertia = engineInertia;
    double fNewEngineInertia = fEngineInertia;
    if (ratio != 0.0) { const double fInertiaSum = D.driveInertia + D.shaftInertiaL + D.shaftInertiaR; fNewEngineInertia = fInertiaSum / (ratio * ratio) + D.clutchInertia + fEngineInertia; }
    const double fInertiaFromWheels = dt_inertia_from_wheels(PP, c, ratio, engineInertia);
    double fDeltaDriveV = 0, fClutchTorq = 0;
    if (!c.clutchOpenState) {
        const double fDeltaRootV = (outTorque / fNewEngineInertia) * dt;
        c.rootVel += fDeltaRootV;
        if (ratio == 0.0) {
            fDeltaDriveV = (tr.feedbackTorque + tl.feedbackTorque) / fInertiaFromWheels * dt;
            c.driveVel += fDeltaDriveV;
        } else {
            const double acc = fDeltaRootV / ratio;
            c.driveVel += acc; c.shaftRVel += acc; c.shaftLVel += acc;
            fDeltaDriveV = (tr.feedbackTorque + tl.feedbackTorque) / fInertiaFromWheels * dt;
            c.rootVel += fDeltaDriveV * ratio;
            c.driveVel += fDeltaDriveV;
        }
    } else {
        fClutchTorq = -((c.engineVel - c.rootVel) / (fabs(c.engineVel - c.rootVel) + 4.0) * (c.locClutch * D.clutchMaxTorque));
        c.currentClutchTorque = fClutchTorq;
        if (ratio != 0.0) {
            c.engineVel += (fClutchTorq + outTorque) / fEngineInertia * dt;
            const double fDeltaRootV = (-fClutchTorq / (fNewEngineInertia - fEngineInertia)) * dt;
            c.rootVel += fDeltaRootV;
            const double acc = fDeltaRootV / ratio;
            c.driveVel += acc; c.shaftRVel += acc; c.shaftLVel += acc;
            fDeltaDriveV = (tr.feedbackTorque + tl.feedbackTorque) / fInertiaFromWheels * dt;
            c.rootVel += fDeltaDriveV * ratio;
            c.driveVel += fDeltaDriveV;
        } else {
            const double fNewEngineVelocity = c.engineVel + outTorque / fEngineInertia * dt;
            c.engineVel = fNewEngineVelocity; c.rootVel = fNewEngineVelocity;
            fDeltaDriveV = (tr.feedbackTorque + tl.feedbackTorque) / fInertiaFromWheels * dt;
            c.driveVel += fDeltaDriveV;
        }
    }
    c.shaftLVel += fDeltaDriveV; c.shaftRVel += fDeltaDriveV;
    if (D.diffType == 1) { c.shaftLVel = c.driveVel; c.shaftRVel = c.driveVel; }
    else {
        double fOutClutchTorq, fDiffLoad;
        if (fClutchTorq != 0.0) fOutClutchTorq = -fClutchTorq; else fOutClutchTorq = c.locClutch * outTorque;
        if (fOutClutchTorq <= 0.0) fDiffLoad = fabs(ratio * D.diffCoastRamp * fOutClutchTorq); else fDiffLoad = fabs(ratio) * (D.diffPowerRamp * fOutClutchTorq);
        const double fDiffTotalLoad = fDiffLoad + D.diffPreLoad;
        if (fabs(c.shaftLVel - c.driveVel) >= 0.1 || fabs(tr.feedbackTorque - tl.feedbackTorque) > fDiffTotalLoad) {
            const double fUnk1 = -((c.shaftLVel - c.shaftRVel) / (fabs(c.shaftLVel - c.shaftRVel) + 0.01) * fDiffTotalLoad);
            const double fDeltaV1 = dt * (fUnk1 / D.shaftInertiaL * 0.5);
            c.shaftLVel += fDeltaV1; c.shaftRVel -= fDeltaV1;
            const double fDeltaV2 = dt * ((tr.feedbackTorque - tl.feedbackTorque) / D.shaftInertiaR * 0.5);
            c.shaftLVel -= fDeltaV2; c.shaftRVel += fDeltaV2;
        } else { c.shaftLVel = c.driveVel; c.shaftRVel = c.driveVel; }
    }
    if (tl.isLocked && tr.isLocked) {
        const float fTorqL = (1.0f * tl.brakeTorque) + tl.handBrakeTorque;
        const float fTorqR = (1.0f * tr.brakeTorque) + tr.handBrakeTorque;
        bool bFlag = true;
        if (fabs(ratio * outTorque) <= (fTorqL + fTorqR)) { if (c.speed <= 1.0f) bFlag = false; }
        if (bFlag) { tl.isLocked = 0; tr.isLocked = 0; }
        else if (c.clutchOpenState) { c.rootVel = 0; c.driveVel = 0; c.shaftLVel = 0; c.shaftRVel = 0; }
    }
    if (!c.clutchOpenState) c.engineVel = c.rootVel;
    tl.angularVelocity = (float)c.shaftLVel; tr.angularVelocity = (float)c.shaftRVel;
    const float fGearTorque = (float)(c.locClutch * outTorque * curGearRatio);
    /* rear rigid axle: torque reaction about the body / axle local z (Drivetrain.cpp:547-563) */
    return fGearTorque * PP.axle.torqueReaction;   /* +z local on the chassis, -z local on the axle */
}

} // namespace pd
