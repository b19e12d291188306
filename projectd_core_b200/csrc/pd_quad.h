/*
 * pd_quad.h -- the tick with FOUR LANES PER CAR (8 cars per warp).
 *
 * Same arithmetic as car_tick (pd_tick.h), distributed over a quad of adjacent lanes:
 *
 *   lane 0  wheel LF : hub0 + strut body 0, solver group "LF strut" (11 rows)
 *   lane 1  wheel RF : hub1 + strut body 1, solver group "RF strut" (11 rows)
 *   lane 2  wheel LR : rigid axle (owner),   solver group "axle"     (5 rows)
 *   lane 3  wheel RR : rigid axle (copy) + fuel tank, solver group "tank" (6 rows)
 *
 *   per-wheel work (suspension, wheel ray cast, tyre forces, 12x3 thermal grid)      -> one wheel per lane
 *   per-group work (constraint rows, 11x11 block LDL^T, back-substitution)           -> one group per lane
 *   track-bound probes                                                              -> probes r and r+4 on lane r
 *   car-level scalar work (controls, assists, engine + drivetrain in fp64, scoring) -> computed redundantly by
 *     all four lanes from identical inputs (it costs the same issue slots as running it on one lane and
 *     removes every broadcast)
 *
 * Lanes exchange data only through the `Ex` policy: sum over the quad (chassis force / torque partials, the
 * 6x6 Schur complement), fetch-from-lane (wheel summaries, partner hub position, probe results) and an
 * all-of vote.  On the GPU these are __shfl_sync with the quad's own 4-bit member mask, so quads of one warp
 * never wait for each other; tests/hostsim runs the four lanes as four host threads.
 * Sums over the quad are a fixed butterfly ((l0+l1)+(l2+l3) on every lane), hence deterministic and
 * independent of how cars are packed into warps or sharded over GPUs.
 */
#pragma once
#include "pd_tick.h"

namespace pd {

PD_HD void lane_body_mass(const PdCarParams& P, int bodyIdx, Body& b, int topo = 0) {
    if (PD_TOPO_FRONT_DW(topo) && (bodyIdx == PD_BODY_HUB0 || bodyIdx == PD_BODY_HUB1)) { const PdDW& D = P.dw[(bodyIdx - PD_BODY_HUB0) >> 1]; b.mass = D.hubMass; b.I = v3(D.hubInertia[0], D.hubInertia[1], D.hubInertia[2]); return; }
    if (PD_TOPO_REAR_DW(topo) && (bodyIdx == PD_BODY_HUB2 || bodyIdx == PD_BODY_HUB3)) { const PdDW& D = P.dw[2 + (bodyIdx - PD_BODY_HUB2)]; b.mass = D.hubMass; b.I = v3(D.hubInertia[0], D.hubInertia[1], D.hubInertia[2]); return; }
    switch (bodyIdx) {
    case PD_BODY_CHASSIS: b.mass = P.chassisMass; b.I = v3(P.chassisInertia[0], P.chassisInertia[1], P.chassisInertia[2]); break;
    case PD_BODY_TANK: b.mass = P.tankMass; b.I = v3(P.tankInertia[0], P.tankInertia[1], P.tankInertia[2]); break;
    case PD_BODY_AXLE: b.mass = P.axle.axleMass; b.I = v3(P.axle.axleInertia[0], P.axle.axleInertia[1], P.axle.axleInertia[2]); break;
    case PD_BODY_HUB0: case PD_BODY_HUB1: { const PdStrut& S = P.strut[(bodyIdx - PD_BODY_HUB0) >> 1]; b.mass = S.hubMass; b.I = v3(S.hubInertia[0], S.hubInertia[1], S.hubInertia[2]); break; }
    default: { const PdStrut& S = P.strut[(bodyIdx - PD_BODY_STRUT0) >> 1]; b.mass = S.strutMass; b.I = v3(S.strutInertia[0], S.strutInertia[1], S.strutInertia[2]); break; }
    }
}

/* TOPO (PD_TOPO_*): with a double-wishbone axle the lane's wheel body is the corner's own hub and its joint group the hub's five
 * links (a single-body group, the code path of the rigid axle); the fuel tank's group stays on lane 3, which then factors two small
 * groups; nothing is sent between the rear lanes (each owns its hub). */
template <int STRIDE, int STRIDE_D, int TOPO = 0, class Ex, class SVX>
PD_HDN void car_tick_quad(const PdCarParams& P, const TrackDev& T, const SVX& sv, float dt, double physicsTime, Ex& ex, float* scratch, float* scratchD, int collPre = -1, volatile uint32_t* collWait = nullptr, const float* cont = nullptr) {
    const int lane = ex.lane;
    const bool front = lane < 2;
    constexpr bool FDW = PD_TOPO_FRONT_DW(TOPO), RDW = PD_TOPO_REAR_DW(TOPO);
    const bool laneDW = front ? FDW : RDW;          /* this lane's corner is a double-wishbone one */
    const bool hasStrutBody = front && !FDW;
    /* Car-level state: with a stride-1 view (the shared-memory staging copy) the four lanes work IN PLACE on the
     * record's car part.  They execute the car-level code converged (ex.sync() below re-joins them after every
     * lane-dependent section) and from identical inputs, so every store is four identical stores and every
     * read-modify-write reads before any lane writes.  (tests/hostsim gives each host thread its own record.) */
    CarS cLocal; CarS* cp = &cLocal;
    if constexpr (sv_traits<SVX>::in_place) cp = car_in_place(sv); else load_car(sv, cLocal);
    CarCtx X(*cp); X.dt = dt; X.time = physicsTime;
    CarS& c = X.c;
#if defined(PD_PHASE_CLOCKS) && defined(__CUDA_ARCH__)
    X.ph = ex.ph;
#endif
    PD_PHASE(X, 0);
    ex.sync();
    Body C, W, S;
    const int wIdx = front ? (PD_BODY_HUB0 + 2 * lane) : (RDW ? PD_BODY_HUB2 + (lane - 2) : PD_BODY_AXLE);
    const int sIdx = hasStrutBody ? (PD_BODY_STRUT0 + 2 * lane) : PD_BODY_TANK;      /* second body: the strut body, or the tank (owned by lane 3, read-only elsewhere) */
    load_body(sv, PD_BODY_CHASSIS, C); lane_body_mass(P, PD_BODY_CHASSIS, C);
    load_body(sv, wIdx, W); lane_body_mass(P, wIdx, W, TOPO);
    load_body(sv, sIdx, S); lane_body_mass(P, sIdx, S, TOPO);

    /* ---------------- Car::step prologue (all lanes, identical) ---------------- */
    c.speed = len(C.v);
    c.collisionFlag = 0; c.outOfTrackFlag = 0;
    {
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
    {
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
    const bool bAllTyresLoaded = ex.all(!(sv.f(PD_OFF_TYRE(lane) + PD_TYRE_o_load) <= 0.0f));
    autoclutch_step(P, X);
    {
        const float fAngVelSq = sqlen(C.w);
        if (c.speed >= 0.5f || fAngVelSq >= 1.0f) c.sleepingFrames = 0;
        else {
            if (bAllTyresLoaded && (c.ctlGas <= 0.01f || c.ctlClutch <= 0.01f || c.currentGear == 1)) c.sleepingFrames++; else c.sleepingFrames = 0;
            if (c.sleepingFrames > P.framesToSleep) { body_stop(C); if (lane == 3) body_stop(S); }
        }
    }
    {
        const V3 vBodyVel = C.v;
        const V3 vAccel = (vBodyVel - v3(c.lastVelX, c.lastVelY, c.lastVelZ)) * (1.0f / dt) * 0.10197838f;
        c.lastVelX = vBodyVel.x; c.lastVelY = vBodyVel.y; c.lastVelZ = vBodyVel.z;
        const V3 g = irot(C.fr, vAccel);
        c.accGX = g.x; c.accGY = g.y; c.accGZ = g.z;
    }
    {
        const float fRpm = engine_rpm(c);
        float heat = 0;
        if (fRpm > (P.engine.minimum * 0.8f)) { const float fLimiter = (float)(int)(P.engine.limiter * P.engine.limiterMultiplier); heat += (((fRpm / fLimiter) * 20.0f) * c.ctlGas) + 85.0f; }
        const float fOneDivMass = 1.0f / P.waterTmass;
        const float fCool = 1.0f - (P.waterCoolSpeedK * c.speed);
        c.waterT += (((((fCool * P.ambientTemperature) - c.waterT) * fOneDivMass) * dt) * P.waterCoolFactor);
        if (heat != 0.0f) c.waterT += ((((heat - c.waterT) * fOneDivMass) * dt) * P.waterHeatFactor);
    }

    /* ---------------- stepComponents: one wheel per lane ---------------- */
    float brakeT[4], handT[4];
    brakes_step(P.brakes, c, sv, P.ambientTemperature, dt, brakeT, handT);      /* all lanes, identical (reads the four tyres' state of the last tick before any lane's tyre step writes) */
    PD_PHASE(X, 1);
    const float myBrake = brakeT[lane], myHand = handT[lane];          /* per wheel: disc temperatures (cars with [TEMPS_*]) scale each wheel's torque */
    float travel, dspeed;
    Frame hf;
    if (laneDW) { dw_step(P.dw[lane], C, W, travel, dspeed); hf = dw_hub_frame(P.dw[lane], W); }
    else if (front) { strut_step(P.strut[lane], C, W, travel, dspeed); hf = strut_hub_frame(P.strut[lane], W); }
    else { axle_step(P.axle, C, W, lane - 2, travel, dspeed); hf = axle_hub_frame(P.axle, W, lane - 2); }
    sv.f(PD_OFF_TYRE(lane) + PD_TYRE_o_suspTravel, travel); sv.f(PD_OFF_TYRE(lane) + PD_TYRE_o_suspDamperSpeed, dspeed);
    WheelLink my;
    PD_PHASE(X, 2);
    tyre_step(P, T, lane, X, sv, W, hf, C, myBrake, myHand, my);
    PD_PHASE(X, 5);
    for (int w = 0; w < 4; ++w) {
        WheelLink& L = X.wl[w];
        L.load = ex.get(my.load, w); L.feedbackTorque = ex.get(my.feedbackTorque, w); L.angularVelocity = ex.get(my.angularVelocity, w);
        L.brakeTorque = ex.get(my.brakeTorque, w); L.handBrakeTorque = ex.get(my.handBrakeTorque, w); L.ndSlip = ex.get(my.ndSlip, w);
        L.slipRatio = ex.get(my.slipRatio, w); L.isLocked = ex.get(my.isLocked, w); L.surfaceId = ex.get(my.surfaceId, w);
    }
    if constexpr (FDW || RDW) if (P.heave[0].present || P.heave[1].present) {      /* heave springs (uniform condition: every lane of the quad takes part in the exchange) */
        const V3 pp = ex.get(W.fr.p, lane ^ 1), pv = ex.get(W.v, lane ^ 1);
        if (laneDW && P.heave[front ? 0 : 1].present && P.heave[front ? 0 : 1].k != 0.0f) {      /* this axle's spring: each lane applies what acts on its own hub */
        const int a = front ? 0 : 2;
        Body dummy = W;
        if ((lane & 1) == 0) heave_step(P.heave[front ? 0 : 1], P.dw[a], P.dw[a + 1], C, W, W.fr.p, W.v, dummy, pp, pv, 0);
        else heave_step(P.heave[front ? 0 : 1], P.dw[a], P.dw[a + 1], C, dummy, pp, pv, W, W.fr.p, W.v, 1);
        }
    }
    if (lane == 0) aero_step(P, C);          /* all wings on one lane: the chassis force is then accumulated in the reference's order (a per-lane split is ~2 % faster but re-associates the sum, and the 1 s free-running divergence test is sensitive to that) */
    V3 steerA1 = v3(0, 0, 0), steerA2 = v3(0, 0, 0);
    if (front) { /* SteeringSystem::step */
        const float steer = -c.finalSteerAngleSignal * P.steerLinearRatio;
        const float* refPoint = FDW ? P.dw[lane].refPoint : P.strut[lane].refPoint;
        const float* baseCarSteer = FDW ? P.dw[lane].baseCarSteer : P.strut[lane].baseCarSteer;
        const float* tyreSteer = FDW ? P.dw[lane].tyreSteer : P.strut[lane].tyreSteer;
        const float sx = signf_(refPoint[0]);
        const float offx = 0.0f + steer + (sx * (FDW ? P.dw[lane].toeOutLinear : P.strut[lane].toeOutLinear));
        const V3 carSteer = v3(baseCarSteer[0] + offx, baseCarSteer[1], baseCarSteer[2]);
        steerA1 = to_local(C.fr, to_world(C.fr, carSteer));
        steerA2 = to_local(W.fr, to_world(W.fr, v3(tyreSteer[0], tyreSteer[1], tyreSteer[2])));
    }
    ex.sync();
    PD_PHASE(X, 6);
    autoblip_step(P, X);
    autoshift_step(P, X);
    gearchanger_step(P, X);
    const float fAxleTorq = drivetrain_step(P, X);
    PD_PHASE(X, 7);
    if (P.tyre[lane].driven) { sv.f(PD_OFF_TYRE(lane) + PD_TYRE_o_angularVelocity, X.wl[lane].angularVelocity); sv.i(PD_OFF_TYRE(lane) + PD_TYRE_o_isLocked, X.wl[lane].isLocked); }
    if (!RDW && lane == 2) { add_rel_torque(C, v3(0, 0, fAxleTorq)); add_rel_torque(W, v3(0, 0, -fAxleTorq)); }       /* Drivetrain.cpp:547-553: only with a rigid rear axle */
    { /* anti-roll bars */
        const V3 partner = ex.get(W.fr.p, lane ^ 1);
        if (front || RDW) {       /* two hubs: each lane applies its own half of AntirollBar::step */
            const float k = P.arbK[front ? 0 : 1];
            if (k > 0.0f) {
                const bool first = (lane & 1) == 0;
                const V3 hubWorld0 = first ? W.fr.p : partner, hubWorld1 = first ? partner : W.fr.p;
                const V3 vHubLoc0 = to_local(C.fr, hubWorld0), vHubLoc1 = to_local(C.fr, hubWorld1);
                const float fDeltaK = (vHubLoc1.y - vHubLoc0.y) * k;
                const V3 vForce = norm(C.fr.ay) * fDeltaK;
                if (first) { add_force_at_pos(W, vForce, hubWorld0); add_rel_force_at_rel_pos(C, v3(0, -fDeltaK, 0), vHubLoc0); }
                else { add_force_at_pos(W, vForce * -1.0f, hubWorld1); add_rel_force_at_rel_pos(C, v3(0, fDeltaK, 0), vHubLoc1); }
            }
        } else if (lane == 2) {
            const Frame f0 = axle_hub_frame(P.axle, W, 0), f1 = axle_hub_frame(P.axle, W, 1);
            arb_step(P.arbK[1], C, W, f0.p, W, f1.p);
        }
    }
    /* chassis force / torque: sum of the four lanes' partials; the axle collects lane 3's wheel forces */
    C.F = ex.sum(C.F); C.T = ex.sum(C.T);
    if constexpr (!RDW) {
        const V3 f3 = ex.get(W.F, 3), t3 = ex.get(W.T, 3);
        if (lane == 2) { W.F += f3; W.T += t3; }
    }

    /* ---------------- collisionStep (odd frames): the cell lists are dealt to the four lanes, any hit sets the flag ---------------- */
    /* three ways to the answer: k_collide ran ahead (collPre 0 / 1); the block's collision warp is working on it right now and
       posts it in shared memory (collWait: picked up just before the scoring, the only consumer); or the quad tests here */
    bool collDeferred = false;
    if (c.physFrame & 1) {
        if (collPre >= 0) { if (collPre != 0) c.collisionFlag = 1; }
        else if (collWait) collDeferred = true;
        else { const bool hit = car_collide(P, T, C, lane, 4); if (!ex.all(!hit)) c.collisionFlag = 1; }
    }
    const bool freshContacts = (c.physFrame & 1) != 0;
    c.physFrame++;
    /* ---------------- dWorldStep: one joint group per lane ---------------- */
    PD_PHASE(X, 8);
    const float h = dt, hinv = 1.0f / dt;
    BodyDyn dC, dA, dB;
    body_dyn(C, P.gravityY, h, dC);
    Body& A = (!RDW && lane == 3) ? S : W;        /* rigid rear axle: lane 3 owns the tank, the other lanes their wheel body; double-wishbone rear: every lane its hub */
    body_dyn(A, P.gravityY, h, dA);
    if (hasStrutBody || (RDW && lane == 3)) body_dyn(S, P.gravityY, h, dB); else dB = dA;       /* dB: the strut body, or (DWB rear, lane 3) the tank */
    float S21[21], b6[6];
    for (int k = 0; k < 21; ++k) S21[k] = 0;
    for (int k = 0; k < 6; ++k) b6[k] = 0;
#if PD_SOLVER2
    /* one group per lane, register-resident (pd_solver2.h): lanes 0 / 1 the strut groups, lanes 2 / 3 the single-body groups
       (axle, tank) through one shared code path */
    float GR[FDW ? 105 : PD_GSYS_WORDS];
    float GT[RDW ? 105 : 1];                  /* DWB rear: the tank's group, on lane 3 beside its hub's */
    StrutSys GS; GS.R = GR; SingleSys G1; G1.R = GR; SingleSys G2; G2.R = GT;
    if (hasStrutBody) strut_factor(P, P.strut[lane], C, W, S, steerA1, steerA2, dA, dB, dC, hinv, X.dballErp, X.dballCfm, GS, S21, b6);
    else {
        float cfm[6];
        if (laneDW) {
            const V3 st[2] = {steerA1, steerA2};
            const bool ml = P.dw[lane].multilink != 0;       /* SuspensionML's joints keep the world's ERP / CFM */
            single_rows_links(P.dw[lane].link, PD_DW_LINKS, C, W, hinv, ml ? P.worldERP : X.dballErp, ml ? P.worldCFM : X.dballCfm, G1, cfm, front ? st : nullptr);
        }
        else if (lane == 2) single_rows_axle(P, C, W, hinv, X.dballErp, X.dballCfm, G1, cfm);
        else single_rows_tank(P, S, C, hinv, G1, cfm);
        single_factor(G1, cfm, dA, dC, hinv, S21, b6);
        if constexpr (RDW) { if (lane == 3) { single_rows_tank(P, S, C, hinv, G2, cfm); single_factor(G2, cfm, dB, dC, hinv, S21, b6); } }
    }
    PD_PHASE(X, 9);
    PD_PHASE(X, 10);
    for (int k = 0; k < 21; ++k) S21[k] = ex.sum(S21[k]);
    for (int k = 0; k < 6; ++k) b6[k] = ex.sum(b6[k]);
    schur_add_chassis(S21, C);
    float z[6];
    if (cont && reinterpret_cast<const int*>(cont)[0] > 0) {      /* live contact joints (rare): every lane of the quad solves the same small system */
        float life = c.lifeLeft;
        float dmg[5] = {c.damageZone0, c.damageZone1, c.damageZone2, c.damageZone3, c.damageZone4};
        contacts_solve(P, cont, C, dC, S21, b6, h, freshContacts, life, dmg, z);
        X.newDamage = fabsf(dmg[0] - c.damageZone0) > 0.001f || fabsf(dmg[1] - c.damageZone1) > 0.001f || fabsf(dmg[2] - c.damageZone2) > 0.001f || fabsf(dmg[3] - c.damageZone3) > 0.001f || fabsf(dmg[4] - c.damageZone4) > 0.001f;
        ex.sync();
        c.lifeLeft = life; c.damageZone0 = dmg[0]; c.damageZone1 = dmg[1]; c.damageZone2 = dmg[2]; c.damageZone3 = dmg[3]; c.damageZone4 = dmg[4];
        ex.sync();
    } else solve6(S21, b6, z);
    PD_PHASE(X, 11);
    float cfA[6], cfB[6];
    if (hasStrutBody) strut_backsolve(GS, z, cfA, cfB);
    else { single_backsolve(G1, z, cfA); if constexpr (RDW) { if (lane == 3) single_backsolve(G2, z, cfB); } }
#else
    GScr<STRIDE, STRIDE_D> G; G.bind(scratch, scratchD);
    if (front) build_strut(P, P.strut[lane], C, W, S, steerA1, steerA2, hinv, X.dballErp, X.dballCfm, G);
    else if (lane == 2) build_axle(P, C, W, hinv, X.dballErp, X.dballCfm, G);
    else build_tank(P, S, C, hinv, G);
    PD_PHASE(X, 9);
    ex.sync();
    build_D(G, dA, dB, dC, hinv, PD_GMAX, true, ex.half, ex.nhalf);     /* a helper quad builds every other row pair */
    ex.sync();
    factor_group(G, dA, dB, dC, hinv, S21, b6, PD_GMAX, true, true);
    PD_PHASE(X, 10);
    for (int k = 0; k < 21; ++k) S21[k] = ex.sum(S21[k]);
    for (int k = 0; k < 6; ++k) b6[k] = ex.sum(b6[k]);
    schur_add_chassis(S21, C);
    float z[6];
    solve6(S21, b6, z);
    PD_PHASE(X, 11);
    float cfA[6], cfB[6];
    backsolve_group(G, z, cfA, cfB);
#endif
    const bool hasB = hasStrutBody || (RDW && lane == 3);        /* a second own body: the strut body / the tank beside a DWB rear hub */
    apply_update(A, dA, cfA, h);
    if (hasB) apply_update(S, dB, cfB, h);
    chassis_update(C, dC, z, h);
    integrate_body(C, h); integrate_body(A, h);
    if (hasB) integrate_body(S, h);
    float nf = body_nonfinite_acc(A) + body_nonfinite_acc(C);
    if (hasB) nf += body_nonfinite_acc(S);
    int bad = (nf == 0.0f) ? 0 : 1;
    bad = ex.all(!bad) ? 0 : 1;
    /* own bodies back to the state */
    if (lane == 0) store_body(sv, PD_BODY_CHASSIS, C);
    if (RDW || lane != 3) store_body(sv, wIdx, W);
    if (hasStrutBody || lane == 3) store_body(sv, sIdx, S);

    /* ---------------- Car::postStep ---------------- */
    ex.sync();
    PD_PHASE(X, 12);
    {
        const V3 bodyPos = C.fr.p;
        const int nFat = T.info.nFatPoints;
        if (P.nProbes > 0 && nFat > 0) {
            const V3 cache = v3(c.pointCacheX, c.pointCacheY, c.pointCacheZ);
            if (sqlen(cache - bodyPos) > 1.0f * 1.0f) { c.pointCacheX = bodyPos.x; c.pointCacheY = bodyPos.y; c.pointCacheZ = bodyPos.z; }
        }
        const V3 cachePos = v3(c.pointCacheX, c.pointCacheY, c.pointCacheZ);
        const float nearR = (P.nProbes > 0) ? P.probeLength[0] : T.info.hashCellSize;
        const float nearRSq = nearR * nearR;
        /* this lane's probes: r = lane and lane + 4 */
        float pr[2] = {FLT_MAX, FLT_MAX};
        float rax = bodyPos.x, raz = bodyPos.z, rbx[2] = {0, 0}, rbz[2] = {0, 0};
        bool mine[2] = {false, false};
        bool needBrute = false;
        for (int k = 0; k < 2; ++k) {
            const int r = lane + 4 * k;
            if (r >= P.nProbes) continue;
            mine[k] = true;
            const V3 rayStart = to_world(C.fr, v3(0, 0, 0));
            const V3 rayEndL = to_world(C.fr, v3(P.probeDir[r][0], P.probeDir[r][1], P.probeDir[r][2]) * P.probeLength[r]);
            const V3 dir = norm(rayEndL - rayStart);
            const V3 rayEnd = rayStart + dir * (P.probeLength[r] * 1.1f);
            rax = rayStart.x; raz = rayStart.z; rbx[k] = rayEnd.x; rbz[k] = rayEnd.z;
            if (ex.nhalf == 2) continue;                            /* with a helper quad the walks happen below, one per twin lane */
            if (!probe_walk(T, rax, raz, rbx[k], rbz[k], cachePos, nearRSq, pr[k])) needBrute = true;
        }
        if (ex.nhalf == 2) {
            /* the main lane walks probe `lane`, its twin probe `lane + 4` -- at the SAME call site, so the two walks run side by side */
            const int kk = ex.half;
            const bool have = kk ? mine[1] : mine[0];
            const float bx = kk ? rbx[1] : rbx[0], bz = kk ? rbz[1] : rbz[0];
            float best = FLT_MAX;
            if (have) { if (!probe_walk(T, rax, raz, bx, bz, cachePos, nearRSq, best)) needBrute = true; }
            if (kk) pr[1] = best; else pr[0] = best;
        }
        if (ex.nhalf == 2) {
            const float o0 = ex.peer(pr[0]), o1 = ex.peer(pr[1]); const int ob = ex.peer(needBrute ? 1 : 0);
            if (ex.half == 1) pr[0] = o0; else pr[1] = o1;
            needBrute = needBrute || ob != 0;
        }
        PD_PHASE(X, 13);
        int bestPoint = 0;
        const bool haveNearest = nearest_point_grid_quad(T, bodyPos, cachePos, nearRSq, ex, bestPoint);
        PD_PHASE(X, 14);
        if (!ex.all(!needBrute && haveNearest)) {
            /* exhaustive form of the reference (car far off the indexed area); every lane scans, each for its own probes */
            float bestDistSq = FLT_MAX; bestPoint = 0; pr[0] = FLT_MAX; pr[1] = FLT_MAX;
            for (int id = 0; id < nFat; ++id) {
                const PdFatPoint& f = T.fat[id];
                const V3 loc = v3(f.best[0], f.best[1], f.best[2]);
                if (!(sqlen(cachePos - loc) < nearRSq)) continue;
                { const float dsq = sqlen(loc - bodyPos); if (bestDistSq > dsq) { bestDistSq = dsq; bestPoint = id; } }
                const PdFatPoint& g = T.fat[id + 1 < nFat ? id + 1 : 0];
                for (int k = 0; k < 2; ++k) {
                    if (!mine[k]) continue;
                    float ix, iz;
                    if (line_intersection(rax, raz, rbx[k], rbz[k], f.left[0], f.left[2], g.left[0], g.left[2], ix, iz)) { const float dx = rax - ix, dz = raz - iz; pr[k] = tminf(pr[k], sqrtf(dx * dx + dz * dz)); }
                    if (line_intersection(rax, raz, rbx[k], rbz[k], f.right[0], f.right[2], g.right[0], g.right[2], ix, iz)) { const float dx = rax - ix, dz = raz - iz; pr[k] = tminf(pr[k], sqrtf(dx * dx + dz * dz)); }
                }
            }
        }
        PD_PHASE(X, 15);
        for (int r = 0; r < P.nProbes && r < 8; ++r) {
            const float v = ex.get((r < 4) ? pr[0] : pr[1], r & 3);
            c.probes[r] = (v != FLT_MAX) ? v : P.probeLength[r];
        }
        if (c.nearestTrackPointId != bestPoint) { c.oldTrackPointId = c.nearestTrackPointId; c.nearestTrackPointId = bestPoint; c.lastTrackPointTimestamp = (float)physicsTime; }
        c.oldTrackLocation = c.trackLocation; c.trackLocation = 0;
        if (bestPoint >= 0 && bestPoint < nFat) {
            if (nFat >= 5) {
                int prevId = bestPoint - 1; if (prevId < 0) prevId = nFat - 1;
                int nextId = bestPoint + 1; if (nextId >= nFat) nextId = 0;
                int prevId2 = prevId - 1; if (prevId2 < 0) prevId2 = nFat - 1;
                int nextId2 = nextId + 1; if (nextId2 >= nFat) nextId2 = 0;
                int sid; float sdist;
                if (spline_nearest_quad(T, bodyPos, prevId2 * T.info.interpolateStep, nextId2 * T.info.interpolateStep, ex, sid, sdist)) {
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
    ex.sync();
    PD_PHASE(X, 16);
    if (collDeferred) {
        uint32_t v = 0;
        for (int spins = 0; spins < (1 << 26) && (v = *collWait) == 0u; ++spins) { }     /* 1 = clear, 2 = contact (posted by the collision warp of this block) */
        ex.sync();
        if (v == 2u) c.collisionFlag = 1;
        if (lane == 0) *collWait = 0u;                 /* the slot is a pad word of the record: back to 0 before it is stored */
        ex.sync();
    }
    post_lookahead_quad(P, T, C, c, ex);
    post_scoring(P, T, C, X, dt);
    c.episodeSteps++; c.thermalPrimed = 1;
    if (bad) c.nanFlag = 1;
    if constexpr (!sv_traits<SVX>::in_place) { if (lane == 0) store_car(sv, c); }
    ex.sync();
    PD_PHASE(X, 17);
}

} // namespace pd
