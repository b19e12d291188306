/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE (oracle only; never linked into the product).
 *
 * C entry points around the reference's own Simulator / Track / Car objects (compiled unmodified
 * from /root/reference/src/ProjectD, see oracle/Makefile) driven exactly the way the reference's
 * Python binding drives them (src/PyProjectD/PyProjectD.cpp:111-137,160-180,219-237,268-363) and the
 * way pyprojectd/projectd_env.py:118-136,157-227 calls that binding.  Used by tests/ (ctypes), by
 * tests/golden/make_golden.py, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
 * reference legs.  It also exports the full per-car state in the pd_state.h record layout and the
 * car parameters in the pd_params.h layout so that the CUDA path can be compared field by field.
 */
#include <time.h>
#include "Sim/Simulator.h"
#include "Sim/Track.h"
#include "Car/CarImpl.h"
#include "ode_restate/ode_core.h"
#include "../include/pd_state.h"
#include "../include/pd_params.h"
#include <cstring>
#include <string>

namespace D {
oder::World* pdref_world(IPhysicsEngine* e);
oder::Body* pdref_body(IRigidBody* rb);
oder::Joint* pdref_joint(IJoint* j);
void pdref_ray_stats(IPhysicsEngine* e, unsigned long long* rays, unsigned long long* tris);
unsigned int pdref_get_frame(IPhysicsEngine* e);
void pdref_set_frame(IPhysicsEngine* e, unsigned int f);
int pdref_engine_contacts(IPhysicsEngine* e, float* out8, int cap);
void pdref_engine_clear_contacts(IPhysicsEngine* e);
void pdref_engine_set_response(IPhysicsEngine* e, int on);
int pdref_body_box(IRigidBody* rb, float* centre3, float* size3);
int pdref_body_mesh(IRigidBody* rb, ITriMesh** mesh, float* offR9, float* off3);
}

using namespace D;

struct RefSim {
    std::shared_ptr<Simulator> sim;
    Car* car = nullptr;
    std::string err;
};

/* "hub on five distance joints": SuspensionDW or SuspensionML (the record and the kernels treat them alike, PdDW::multilink tells them apart) */
static bool front_dw(Car* c) { return c->suspensionTypeF == SuspensionType::DoubleWishbone || c->suspensionTypeF == SuspensionType::Multilink; }
static bool rear_dw(Car* c) { return c->suspensionTypeR == SuspensionType::DoubleWishbone || c->suspensionTypeR == SuspensionType::Multilink; }
static IRigidBody* dw_hub(ISuspension* s) { return s->getType() == SuspensionType::Multilink ? static_cast<SuspensionML*>(s)->hub.get() : static_cast<SuspensionDW*>(s)->hub.get(); }
/* the body kept in slot b of the record (include/pd_state.h), or null when this car's topology has none there */
static oder::Body* body_of(RefSim* h, int b) {
    Car* c = h->car;
    switch (b) {
    case PD_BODY_CHASSIS: return pdref_body(c->body.get());
    case PD_BODY_TANK: return pdref_body(c->fuelTankBody.get());
    case PD_BODY_HUB0: case PD_BODY_HUB1: {
        ISuspension* s = c->suspensions[(b - PD_BODY_HUB0) / 2];
        return front_dw(c) ? pdref_body(dw_hub(s)) : pdref_body(static_cast<SuspensionStrut*>(s)->hub.get());
    }
    case PD_BODY_STRUT0: case PD_BODY_STRUT1:
        return front_dw(c) ? nullptr : pdref_body(static_cast<SuspensionStrut*>(c->suspensions[(b - PD_BODY_STRUT0) / 2])->strutBody.get());
    case PD_BODY_AXLE: return rear_dw(c) ? pdref_body(dw_hub(c->suspensions[2])) : pdref_body(c->rigidAxle.get());
    case PD_BODY_HUB3: return rear_dw(c) ? pdref_body(dw_hub(c->suspensions[3])) : nullptr;
    }
    return nullptr;
}

/* ---- record packing helpers ---- */
static inline void putF(uint32_t* r, int o, float v) { memcpy(r + o, &v, 4); }
static inline void putI(uint32_t* r, int o, int32_t v) { memcpy(r + o, &v, 4); }
static inline void putD(uint32_t* r, int o, double v) { memcpy(r + o, &v, 8); } /* little endian: lo, hi */
static inline float getF(const uint32_t* r, int o) { float v; memcpy(&v, r + o, 4); return v; }
static inline int32_t getI(const uint32_t* r, int o) { int32_t v; memcpy(&v, r + o, 4); return v; }
static inline double getD(const uint32_t* r, int o) { double v; memcpy(&v, r + o, 8); return v; }

static int surface_index(Track* t, Surface* s) {
    if (!s) return -1;
    for (size_t i = 0; i < t->surfaces.size(); ++i) if (t->surfaces[i].get() == s) return (int)i;
    return -1;
}

static void copy_curve(PdCurve& d, const Curve& c) {
    memset(&d, 0, sizeof(d));
    d.n = c.getCount();
    if (d.n > PD_CURVE_MAX) { fprintf(stderr, "[oracle] curve too long %d\n", d.n); d.n = PD_CURVE_MAX; }
    for (int i = 0; i < d.n; ++i) { d.ref[i] = c.references[i]; d.val[i] = c.values[i]; }
}
static void copy_damper(PdDamper& d, const Damper& s) {
    d.bumpSlow = s.bumpSlow; d.reboundSlow = s.reboundSlow; d.bumpFast = s.bumpFast; d.reboundFast = s.reboundFast;
    d.fastThresholdBump = s.fastThresholdBump; d.fastThresholdRebound = s.fastThresholdRebound;
}
static void copy_dball(PdDBall& d, IJoint* j, float distance) {
    oder::Joint* oj = pdref_joint(j);
    for (int k = 0; k < 3; ++k) { d.anchor1[k] = oj->anchor1[k]; d.anchor2[k] = oj->anchor2[k]; }
    d.distance = distance;
}
static void v3(float* d, const vec3f& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }

extern "C" {

int pdref_state_words() { return PD_STATE_WORDS; }
int pdref_params_bytes() { return (int)sizeof(PdCarParams); }

void* pdref_create(const char* base_path, const char* track, const char* car_model) {
    RefSim* h = new RefSim();
    try {
        h->sim = std::make_shared<Simulator>();
        h->sim->simulatorId = 0;
        h->sim->init(strw(std::string(base_path)));
        h->sim->loadTrack(strw(std::string(track)));
        h->car = h->sim->addCar(strw(std::string(car_model)));
    } catch (const std::exception& ex) {
        fprintf(stderr, "[oracle] pdref_create failed: %s\n", ex.what());
        delete h; return nullptr;
    }
    return h;
}
void pdref_destroy(void* hv) { delete (RefSim*)hv; }

/* PyProjectD.cpp:268-326 equivalents */
void pdref_teleport_spline(void* hv, float distNorm) { ((RefSim*)hv)->car->teleportToSpline(distNorm); }
void pdref_teleport_mode(void* hv, int mode) { ((RefSim*)hv)->car->teleportByMode((TeleportMode)mode); }
void pdref_set_seed(unsigned int seed) { srand(seed); }
void pdref_set_auto_teleport(void* hv, int onCollision, int onBadLoc, int mode) {
    Car* c = ((RefSim*)hv)->car; c->teleportOnCollision = onCollision; c->teleportOnBadLocation = onBadLoc; c->teleportMode = mode;
}
void pdref_set_assists(void* hv, int autoClutch, int autoShift, int autoBlip) {
    Car* c = ((RefSim*)hv)->car;
    c->autoClutch->useAutoOnStart = autoClutch != 0; c->autoClutch->useAutoOnChange = autoClutch != 0;
    c->autoShift->isActive = autoShift != 0; c->autoBlip->isActive = autoBlip != 0;
}
void pdref_set_tune(void* hv, const char* name, float value) { ((RefSim*)hv)->car->setup->setTune(name, value); }
void pdref_set_scoring_var(void* hv, const char* name, float value) { ((RefSim*)hv)->car->scoring->config->setVar(name, value); }
void pdref_set_controls(void* hv, float steer, float clutch, float brake, float handBrake, float gas,
                        int requestedGear, int gearUp, int gearDn, int smooth) {
    Car* c = ((RefSim*)hv)->car;
    CarControls k;
    k.steer = steer; k.clutch = clutch; k.brake = brake; k.handBrake = handBrake; k.gas = gas;
    k.isShifterSupported = c->controls.isShifterSupported;
    k.requestedGearIndex = (int8_t)requestedGear; k.gearUp = (int8_t)gearUp; k.gearDn = (int8_t)gearDn;
    c->controls = k; c->smoothSteer = smooth;
}
/* PyProjectD.cpp:160-180 */
void pdref_step(void* hv, double dt) {
    Simulator* s = ((RefSim*)hv)->sim.get();
    s->step((float)dt, s->physicsTime, s->gameTime);
    s->physicsTime += dt; s->gameTime += dt;
}
double pdref_get_time(void* hv) { return ((RefSim*)hv)->sim->physicsTime; }
void pdref_set_time(void* hv, double t) { ((RefSim*)hv)->sim->physicsTime = t; ((RefSim*)hv)->sim->gameTime = t; }
void pdref_get_car_state(void* hv, void* out) { memcpy(out, ((RefSim*)hv)->car->state.get(), sizeof(CarState)); }
int pdref_car_state_bytes() { return (int)sizeof(CarState); }

/* ------------------------------------------------------------------------------------------------ */
void pdref_get_state(void* hv, uint32_t* r) {
    RefSim* h = (RefSim*)hv; Car* c = h->car; Track* trk = c->track;
    memset(r, 0, sizeof(uint32_t) * PD_STATE_WORDS);
    for (int b = 0; b < PD_NUM_BODIES; ++b) {
        oder::Body* ob = body_of(h, b); int o = PD_OFF_BODY(b);
        if (!ob) continue;
        putF(r, o + PD_BODY_o_px, ob->pos[0]); putF(r, o + PD_BODY_o_py, ob->pos[1]); putF(r, o + PD_BODY_o_pz, ob->pos[2]);
        putF(r, o + PD_BODY_o_qw, ob->q[0]); putF(r, o + PD_BODY_o_qx, ob->q[1]); putF(r, o + PD_BODY_o_qy, ob->q[2]); putF(r, o + PD_BODY_o_qz, ob->q[3]);
        putF(r, o + PD_BODY_o_vx, ob->lvel[0]); putF(r, o + PD_BODY_o_vy, ob->lvel[1]); putF(r, o + PD_BODY_o_vz, ob->lvel[2]);
        putF(r, o + PD_BODY_o_wx, ob->avel[0]); putF(r, o + PD_BODY_o_wy, ob->avel[1]); putF(r, o + PD_BODY_o_wz, ob->avel[2]);
        /* axes = columns of ODE's row-major R */
        putF(r, o + PD_BODY_o_axx, ob->R[0]); putF(r, o + PD_BODY_o_axy, ob->R[3]); putF(r, o + PD_BODY_o_axz, ob->R[6]);
        putF(r, o + PD_BODY_o_ayx, ob->R[1]); putF(r, o + PD_BODY_o_ayy, ob->R[4]); putF(r, o + PD_BODY_o_ayz, ob->R[7]);
        putF(r, o + PD_BODY_o_azx, ob->R[2]); putF(r, o + PD_BODY_o_azy, ob->R[5]); putF(r, o + PD_BODY_o_azz, ob->R[8]);
    }
    for (int w = 0; w < 4; ++w) {
        Tyre* t = c->tyres[w].get(); const TyreStatus& s = t->status; int o = PD_OFF_TYRE(w);
#define TF(n, v) putF(r, o + PD_TYRE_o_##n, (v))
#define TI(n, v) putI(r, o + PD_TYRE_o_##n, (v))
#define TD(n, v) putD(r, o + PD_TYRE_o_##n, (v))
        TF(depth, s.depth); TF(load, s.load); TF(camberRAD, s.camberRAD); TF(slipAngleRAD, s.slipAngleRAD); TF(slipRatio, s.slipRatio);
        TF(angularVelocity, s.angularVelocity); TF(Fy, s.Fy); TF(Fx, s.Fx); TF(Mz, s.Mz); TI(isLocked, s.isLocked ? 1 : 0);
        TF(slipFactor, s.slipFactor); TF(ndSlip, s.ndSlip); TF(distToGround, s.distToGround); TF(Dy, s.Dy); TF(Dx, s.Dx); TF(D, s.D);
        TF(dirtyLevel, s.dirtyLevel); TF(rollingResistence, s.rollingResistence); TF(thermalInput, s.thermalInput); TF(feedbackTorque, s.feedbackTorque);
        TF(loadedRadius, s.loadedRadius); TF(effectiveRadius, s.effectiveRadius); TF(liveRadius, s.liveRadius);
        TF(pressureStatic, s.pressureStatic); TF(pressureDynamic, s.pressureDynamic); TD(virtualKM, s.virtualKM); TF(inflation, s.inflation);
        TD(flatSpot, s.flatSpot); TF(wearMult, s.wearMult);
        TF(oldAngularVelocity, t->oldAngularVelocity); TF(localMX, t->localMX);
        TF(contactX, t->contactPoint.x); TF(contactY, t->contactPoint.y); TF(contactZ, t->contactPoint.z);
        TF(normalX, t->contactNormal.x); TF(normalY, t->contactNormal.y); TF(normalZ, t->contactNormal.z);
        TI(surfaceId, surface_index(trk, t->surfaceDef)); TI(hasContact, t->surfaceDef ? 1 : 0);
        TF(totalHubVelocity, t->totalHubVelocity); TF(slidingVelocityX, t->slidingVelocityX); TF(slidingVelocityY, t->slidingVelocityY);
        TF(brakeTorque, t->inputs.brakeTorque); TF(handBrakeTorque, t->inputs.handBrakeTorque);
        auto st = c->suspensions[w]->getStatus();
        TF(suspTravel, st.travel); TF(suspDamperSpeed, st.damperSpeedMS);
        TF(coreTemp, t->thermalModel->coreTemp); TD(phase, t->thermalModel->phase);
        TF(practicalTemp, t->thermalModel->practicalTemp); TF(thermalMultD, t->thermalModel->thermalMultD);
        for (int p = 0; p < PD_THERMAL_PATCHES; ++p) putF(r, PD_OFF_TYRE_PATCH(w) + p, t->thermalModel->patches[p].T);
#undef TF
#undef TI
#undef TD
    }
    {
        int o = PD_OFF_CAR; Drivetrain* d = c->drivetrain.get(); Engine* e = d->engineModel.get(); ScoringSystem* sc = c->scoring.get();
#define CF(n, v) putF(r, o + PD_CAR_o_##n, (v))
#define CI(n, v) putI(r, o + PD_CAR_o_##n, (v))
#define CD(n, v) putD(r, o + PD_CAR_o_##n, (v))
        CF(ctlSteer, c->controls.steer); CF(ctlClutch, c->controls.clutch); CF(ctlBrake, c->controls.brake);
        CF(ctlHandBrake, c->controls.handBrake); CF(ctlGas, c->controls.gas);
        CI(ctlRequestedGear, c->controls.requestedGearIndex); CI(ctlGearUp, c->controls.gearUp); CI(ctlGearDn, c->controls.gearDn);
        CI(smoothSteer, c->smoothSteer);
        CF(smoothSteerValue, c->smoothSteerValue); CF(finalSteerAngleSignal, c->finalSteerAngleSignal);
        CF(lastVelX, c->lastVelocity.x); CF(lastVelY, c->lastVelocity.y); CF(lastVelZ, c->lastVelocity.z);
        CF(accGX, c->accG.x); CF(accGY, c->accG.y); CF(accGZ, c->accG.z);
        CD(fuel, c->fuel); CI(sleepingFrames, c->sleepingFrames); CF(waterT, c->water->t); CF(speed, c->speed.value);
        CI(collisionFlag, c->collisionFlag ? 1 : 0); CI(outOfTrackFlag, c->outOfTrackFlag ? 1 : 0);
        CI(nearestTrackPointId, c->nearestTrackPointId); CI(oldTrackPointId, c->oldTrackPointId); CI(splinePointId, c->splinePointId);
        CF(lastTrackPointTimestamp, c->lastTrackPointTimestamp); CF(trackLocation, c->trackLocation); CF(oldTrackLocation, c->oldTrackLocation);
        CF(bodyVsTrack, c->bodyVsTrack); CF(velocityVsTrack, c->velocityVsTrack);
        CF(pointCacheX, trk->pointCachePos.x); CF(pointCacheY, trk->pointCachePos.y); CF(pointCacheZ, trk->pointCachePos.z);
        AutoClutch* ac = c->autoClutch.get();
        CF(acSeqTime, ac->clutchSequence.currentTime); CI(acSeqDone, ac->clutchSequence.isDone ? 1 : 0);
        {
            int prof = 0;
            if (ac->clutchSequence.clutchCurve.getCount() == 4) {
                /* identify which profile the running sequence was copied from (AutoClutch.cpp:202-221) */
                if (ac->downshiftProfile.getCount() == 4 && ac->clutchSequence.clutchCurve.references == ac->downshiftProfile.references) prof = 2;
                else if (ac->upshiftProfile.getCount() == 4 && ac->clutchSequence.clutchCurve.references == ac->upshiftProfile.references) prof = 1;
            }
            CI(acSeqProfile, prof);
        }
        CF(acClutchValueSignal, ac->clutchValueSignal);
        CD(blipStartTime, c->autoBlip->blipStartTime); CF(gasCutoff, c->autoShift->gasCutoff);
        CI(lastGearUp, c->gearChanger->lastGearUp ? 1 : 0); CI(lastGearDn, c->gearChanger->lastGearDn ? 1 : 0);
        CI(reqRequest, (int)d->gearRequest.request); CD(reqTimeAcc, d->gearRequest.timeAccumulator); CD(reqTimeout, d->gearRequest.timeout);
        CI(reqGear, d->gearRequest.requestedGear);
        CD(engineVel, d->engine.velocity); CD(driveVel, d->drive.velocity); CD(shaftLVel, d->outShaftL.velocity); CD(shaftRVel, d->outShaftR.velocity);
        CD(rootVel, d->rootVelocity); CD(locClutch, d->locClutch); CD(lastRatio, d->lastRatio); CD(cutOff, d->cutOff);
        CI(currentGear, d->currentGear); CI(isGearGrinding, d->isGearGrinding ? 1 : 0); CI(clutchOpenState, d->clutchOpenState ? 1 : 0);
        CD(validShiftRPMWindow, d->validShiftRPMWindow); CD(currentClutchTorque, d->currentClutchTorque);
        CI(limiterOn, e->limiterOn); CF(lifeLeft, e->lifeLeft); CF(fuelPressure, e->fuelPressure); CF(gasUsage, e->gasUsage); CD(outTorque, e->status.outTorque);
        CI(drifting, sc->drifting ? 1 : 0); CI(driftExtreme, sc->driftExtreme ? 1 : 0); CI(driftInvalid, sc->driftInvalid ? 1 : 0);
        CF(currentDriftAngle, sc->currentDriftAngle); CF(currentSpeedMultiplier, sc->currentSpeedMultiplier); CF(lastDriftDirection, sc->lastDriftDirection);
        CF(driftStraightTimer, sc->driftStraightTimer); CF(instantDriftDelta, sc->instantDriftDelta); CF(instantDrift, sc->instantDrift); CF(driftPoints, sc->driftPoints);
        CI(driftComboCounter, sc->driftComboCounter); CF(stepReward, sc->stepReward); CF(totalReward, sc->totalReward); CF(prevEpisodeReward, sc->prevEpisodeReward);
        CI(oldPointId, sc->oldPointId); CI(oldSplinePointId, sc->oldSplinePointId);
        CI(episodeSteps, 0); CI(nanFlag, 0);
        CI(thermalPrimed, c->tyres[0]->thermalModel->patches[5].inputT == 0.0f ? 1 : 0);
        CI(physFrame, (int)pdref_get_frame(h->sim->physics.get()));
        { const auto& tb = e->turbos; if (tb.size() > 0) CF(turboRot0, tb[0]->rotation); if (tb.size() > 1) CF(turboRot1, tb[1]->rotation); if (tb.size() > 2) CF(turboRot2, tb[2]->rotation); CF(turboBoost, e->status.turboBoost); }
        { BrakeSystem* bs = c->brakeSystem.get(); CF(brakeDiscT0, bs->discs[0].t); CF(brakeDiscT1, bs->discs[1].t); CF(brakeDiscT2, bs->discs[2].t); CF(brakeDiscT3, bs->discs[3].t); }
        CF(damageZone0, c->damageZoneLevel[0]); CF(damageZone1, c->damageZoneLevel[1]); CF(damageZone2, c->damageZoneLevel[2]); CF(damageZone3, c->damageZoneLevel[3]); CF(damageZone4, c->damageZoneLevel[4]);
        for (size_t i = 0; i < c->probeHits.size() && i < PD_MAX_PROBES; ++i) putF(r, PD_OFF_PROBES + (int)i, c->probeHits[i]);
        for (size_t i = 0; i < c->lookAhead.size() && i < PD_LOOKAHEAD; ++i) putF(r, PD_OFF_LOOKAHEAD + (int)i, c->lookAhead[i]);
#undef CF
#undef CI
#undef CD
    }
}

/* Inverse of pdref_get_state: overwrite the live reference objects with a record ("identical states in"). */
void pdref_set_state(void* hv, const uint32_t* r) {
    RefSim* h = (RefSim*)hv; Car* c = h->car; Track* trk = c->track;
    pdref_engine_clear_contacts(h->sim->physics.get());     /* contact joints are not part of the record: a restored state has none alive */
    for (int b = 0; b < PD_NUM_BODIES; ++b) {
        oder::Body* ob = body_of(h, b); int o = PD_OFF_BODY(b);
        if (!ob) continue;
        for (int k = 0; k < 3; ++k) ob->pos[k] = getF(r, o + PD_BODY_o_px + k);
        for (int k = 0; k < 4; ++k) ob->q[k] = getF(r, o + PD_BODY_o_qw + k);
        ob->R[0] = getF(r, o + PD_BODY_o_axx); ob->R[3] = getF(r, o + PD_BODY_o_axy); ob->R[6] = getF(r, o + PD_BODY_o_axz);
        ob->R[1] = getF(r, o + PD_BODY_o_ayx); ob->R[4] = getF(r, o + PD_BODY_o_ayy); ob->R[7] = getF(r, o + PD_BODY_o_ayz);
        ob->R[2] = getF(r, o + PD_BODY_o_azx); ob->R[5] = getF(r, o + PD_BODY_o_azy); ob->R[8] = getF(r, o + PD_BODY_o_azz);
        for (int k = 0; k < 3; ++k) { ob->lvel[k] = getF(r, o + PD_BODY_o_vx + k); ob->avel[k] = getF(r, o + PD_BODY_o_wx + k); ob->facc[k] = 0; ob->tacc[k] = 0; }
    }
    for (int w = 0; w < 4; ++w) {
        Tyre* t = c->tyres[w].get(); TyreStatus& s = t->status; int o = PD_OFF_TYRE(w);
#define TF(n) getF(r, o + PD_TYRE_o_##n)
#define TI(n) getI(r, o + PD_TYRE_o_##n)
#define TD(n) getD(r, o + PD_TYRE_o_##n)
        s.depth = TF(depth); s.load = TF(load); s.camberRAD = TF(camberRAD); s.slipAngleRAD = TF(slipAngleRAD); s.slipRatio = TF(slipRatio);
        s.angularVelocity = TF(angularVelocity); s.Fy = TF(Fy); s.Fx = TF(Fx); s.Mz = TF(Mz); s.isLocked = TI(isLocked) != 0;
        s.slipFactor = TF(slipFactor); s.ndSlip = TF(ndSlip); s.distToGround = TF(distToGround); s.Dy = TF(Dy); s.Dx = TF(Dx); s.D = TF(D);
        s.dirtyLevel = TF(dirtyLevel); s.rollingResistence = TF(rollingResistence); s.thermalInput = TF(thermalInput); s.feedbackTorque = TF(feedbackTorque);
        s.loadedRadius = TF(loadedRadius); s.effectiveRadius = TF(effectiveRadius); s.liveRadius = TF(liveRadius);
        s.pressureStatic = TF(pressureStatic); s.pressureDynamic = TF(pressureDynamic); s.virtualKM = TD(virtualKM); s.inflation = TF(inflation);
        s.flatSpot = TD(flatSpot); s.wearMult = TF(wearMult);
        t->oldAngularVelocity = TF(oldAngularVelocity); t->localMX = TF(localMX);
        t->contactPoint = vec3f(TF(contactX), TF(contactY), TF(contactZ));
        t->contactNormal = vec3f(TF(normalX), TF(normalY), TF(normalZ));
        { int si = TI(surfaceId); t->surfaceDef = (si >= 0 && si < (int)trk->surfaces.size()) ? trk->surfaces[si].get() : nullptr; }
        t->totalHubVelocity = TF(totalHubVelocity); t->slidingVelocityX = TF(slidingVelocityX); t->slidingVelocityY = TF(slidingVelocityY);
        t->inputs.brakeTorque = TF(brakeTorque); t->inputs.handBrakeTorque = TF(handBrakeTorque);
        t->thermalModel->coreTemp = TF(coreTemp); t->thermalModel->phase = TD(phase);
        t->thermalModel->practicalTemp = TF(practicalTemp); t->thermalModel->thermalMultD = TF(thermalMultD);
        t->thermalModel->coreTInput = 0;
        for (int p = 0; p < PD_THERMAL_PATCHES; ++p) { t->thermalModel->patches[p].T = getF(r, PD_OFF_TYRE_PATCH(w) + p); t->thermalModel->patches[p].inputT = getI(r, PD_OFF_CAR + PD_CAR_o_thermalPrimed) ? 0.0f : h->sim->ambientTemperature; }
#undef TF
#undef TI
#undef TD
    }
    {
        int o = PD_OFF_CAR; Drivetrain* d = c->drivetrain.get(); Engine* e = d->engineModel.get(); ScoringSystem* sc = c->scoring.get();
#define CF(n) getF(r, o + PD_CAR_o_##n)
#define CI(n) getI(r, o + PD_CAR_o_##n)
#define CD(n) getD(r, o + PD_CAR_o_##n)
        c->controls.steer = CF(ctlSteer); c->controls.clutch = CF(ctlClutch); c->controls.brake = CF(ctlBrake);
        c->controls.handBrake = CF(ctlHandBrake); c->controls.gas = CF(ctlGas);
        c->controls.requestedGearIndex = (int8_t)CI(ctlRequestedGear); c->controls.gearUp = (int8_t)CI(ctlGearUp); c->controls.gearDn = (int8_t)CI(ctlGearDn);
        c->smoothSteer = CI(smoothSteer);
        c->smoothSteerValue = CF(smoothSteerValue); c->finalSteerAngleSignal = CF(finalSteerAngleSignal);
        c->lastVelocity = vec3f(CF(lastVelX), CF(lastVelY), CF(lastVelZ)); c->accG = vec3f(CF(accGX), CF(accGY), CF(accGZ));
        c->fuel = CD(fuel); c->sleepingFrames = CI(sleepingFrames); c->water->t = CF(waterT); c->speed.value = CF(speed);
        c->collisionFlag = CI(collisionFlag) != 0; c->outOfTrackFlag = CI(outOfTrackFlag) != 0;
        pdref_set_frame(h->sim->physics.get(), (unsigned int)CI(physFrame));
        { auto& tb = e->turbos; if (tb.size() > 0) tb[0]->rotation = CF(turboRot0); if (tb.size() > 1) tb[1]->rotation = CF(turboRot1); if (tb.size() > 2) tb[2]->rotation = CF(turboRot2); e->status.turboBoost = CF(turboBoost); }
        { BrakeSystem* bs = c->brakeSystem.get(); bs->discs[0].t = CF(brakeDiscT0); bs->discs[1].t = CF(brakeDiscT1); bs->discs[2].t = CF(brakeDiscT2); bs->discs[3].t = CF(brakeDiscT3); }
        c->damageZoneLevel[0] = CF(damageZone0); c->damageZoneLevel[1] = CF(damageZone1); c->damageZoneLevel[2] = CF(damageZone2); c->damageZoneLevel[3] = CF(damageZone3); c->damageZoneLevel[4] = CF(damageZone4);
        for (int i = 0; i < 5; ++i) c->oldDamageZoneLevel[i] = c->damageZoneLevel[i];        /* Car::postStep leaves them equal (Car.cpp:706-707) */
        c->nearestTrackPointId = CI(nearestTrackPointId); c->oldTrackPointId = CI(oldTrackPointId); c->splinePointId = CI(splinePointId);
        c->lastTrackPointTimestamp = CF(lastTrackPointTimestamp); c->trackLocation = CF(trackLocation); c->oldTrackLocation = CF(oldTrackLocation);
        c->bodyVsTrack = CF(bodyVsTrack); c->velocityVsTrack = CF(velocityVsTrack);
        trk->pointCachePos = vec3f(CF(pointCacheX), CF(pointCacheY), CF(pointCacheZ));
        trk->nearbyPoints.clear();
        /* the cache is always refreshed by the first probe (Car.cpp:726-731 -> Track.cpp:505-510) */
        trk->fatPointsHash.queryNeighbours(trk->pointCachePos, trk->nearbyPoints, c->probes.empty() ? 0.0f : c->probes[0].length);
        AutoClutch* ac = c->autoClutch.get();
        ac->clutchSequence.currentTime = CF(acSeqTime); ac->clutchSequence.isDone = CI(acSeqDone) != 0;
        { int prof = CI(acSeqProfile);
          if (prof == 2) ac->clutchSequence.clutchCurve = ac->downshiftProfile; else if (prof == 1) ac->clutchSequence.clutchCurve = ac->upshiftProfile; else ac->clutchSequence.clutchCurve.reset(); }
        ac->clutchValueSignal = CF(acClutchValueSignal);
        c->autoBlip->blipStartTime = CD(blipStartTime); c->autoShift->gasCutoff = CF(gasCutoff);
        c->gearChanger->lastGearUp = CI(lastGearUp) != 0; c->gearChanger->lastGearDn = CI(lastGearDn) != 0;
        d->gearRequest.request = (GearChangeRequest)CI(reqRequest); d->gearRequest.timeAccumulator = CD(reqTimeAcc); d->gearRequest.timeout = CD(reqTimeout);
        d->gearRequest.requestedGear = CI(reqGear);
        d->engine.velocity = CD(engineVel); d->drive.velocity = CD(driveVel); d->outShaftL.velocity = CD(shaftLVel); d->outShaftR.velocity = CD(shaftRVel);
        d->rootVelocity = CD(rootVel); d->locClutch = CD(locClutch); d->lastRatio = CD(lastRatio); d->cutOff = CD(cutOff);
        d->currentGear = CI(currentGear); d->isGearGrinding = CI(isGearGrinding) != 0; d->clutchOpenState = CI(clutchOpenState) != 0;
        d->validShiftRPMWindow = CD(validShiftRPMWindow); d->currentClutchTorque = CD(currentClutchTorque);
        e->limiterOn = CI(limiterOn); e->lifeLeft = CF(lifeLeft); e->fuelPressure = CF(fuelPressure); e->gasUsage = CF(gasUsage); e->status.outTorque = CD(outTorque);
        sc->drifting = CI(drifting) != 0; sc->driftExtreme = CI(driftExtreme) != 0; sc->driftInvalid = CI(driftInvalid) != 0;
        sc->currentDriftAngle = CF(currentDriftAngle); sc->currentSpeedMultiplier = CF(currentSpeedMultiplier); sc->lastDriftDirection = CF(lastDriftDirection);
        sc->driftStraightTimer = CF(driftStraightTimer); sc->instantDriftDelta = CF(instantDriftDelta); sc->instantDrift = CF(instantDrift); sc->driftPoints = CF(driftPoints);
        sc->driftComboCounter = CI(driftComboCounter); sc->stepReward = CF(stepReward); sc->totalReward = CF(totalReward); sc->prevEpisodeReward = CF(prevEpisodeReward);
        sc->oldPointId = CI(oldPointId); sc->oldSplinePointId = CI(oldSplinePointId);
        c->probeHits.resize(c->probes.size());
        for (size_t i = 0; i < c->probeHits.size() && i < PD_MAX_PROBES; ++i) c->probeHits[i] = getF(r, PD_OFF_PROBES + (int)i);
        for (size_t i = 0; i < c->lookAhead.size() && i < PD_LOOKAHEAD; ++i) c->lookAhead[i] = getF(r, PD_OFF_LOOKAHEAD + (int)i);
#undef CF
#undef CI
#undef CD
    }
}

/* ------------------------------------------------------------------------------------------------ */
void pdref_get_params(void* hv, PdCarParams* P) {
    RefSim* h = (RefSim*)hv; Car* c = h->car; Simulator* sim = h->sim.get();
    memset(P, 0, sizeof(*P));
    P->mass = c->mass;
    { oder::Body* b = pdref_body(c->body.get()); P->chassisMass = b->mass; for (int k = 0; k < 3; ++k) P->chassisInertia[k] = b->I[k]; }
    { oder::Body* b = pdref_body(c->fuelTankBody.get()); P->tankMass = b->mass; for (int k = 0; k < 3; ++k) P->tankInertia[k] = b->I[k]; }
    v3(P->fuelTankPos, c->fuelTankPos);
    P->steerLock = c->steerLock; P->steerRatio = c->steerRatio; P->steerLinearRatio = c->steeringSystem->linearRatio;
    P->fuelKG = c->fuelKG; P->fuelConsumptionK = c->fuelConsumptionK; P->maxFuel = c->maxFuel; P->requestedFuel = c->requestedFuel;
    P->framesToSleep = c->framesToSleep;
    P->waterTmass = c->water->tmass; P->waterCoolSpeedK = c->water->coolSpeedK; P->waterCoolFactor = c->water->coolFactor; P->waterHeatFactor = c->water->heatFactor;
    P->baseCarHeight = c->getBaseCarHeight();
    { oder::Joint* j = pdref_joint(c->fuelTankJoint.get()); for (int k = 0; k < 3; ++k) P->tankOffset[k] = j->offset[k]; for (int k = 0; k < 4; ++k) P->tankQrel[k] = j->qrel[k]; }
    P->topology = (front_dw(c) ? 2 : 0) + (rear_dw(c) ? 1 : 0);
    for (int i = 0; i < 4; ++i) {
        if (!(i < 2 ? front_dw(c) : rear_dw(c))) continue;
        if (c->suspensions[i]->getType() == SuspensionType::Multilink) {
            SuspensionML* s = static_cast<SuspensionML*>(c->suspensions[i]); PdDW& d = P->dw[i];
            d.multilink = 1;
            v3(d.refPoint, s->basePosition); v3(d.baseCarSteer, s->baseCarSteerPosition); v3(d.tyreSteer, s->joints[4].ballTyre.relToTyre);
            d.rodLength = s->rodLength; d.k = s->k; d.progressiveK = s->progressiveK; d.packerRange = s->packerRange; d.bumpStopRate = s->bumpStopRate;
            d.bumpStopProgressive = s->bumpStopProgressive; d.bumpStopUp = s->bumpStopUp; d.bumpStopDn = s->bumpStopDn;
            d.toeOutLinear = s->toeOUT_Linear; d.staticCamber = s->staticCamber; d.baseCFM = s->baseCFM;
            copy_damper(d.damper, s->damper);
            for (int l = 0; l < PD_DW_LINKS; ++l) copy_dball(d.link[l], s->joints[l].joint.get(), pdref_joint(s->joints[l].joint.get())->targetDistance);
            { oder::Body* b = pdref_body(s->hub.get()); d.hubMass = b->mass; for (int k = 0; k < 3; ++k) d.hubInertia[k] = b->I[k]; }
            continue;
        }
        SuspensionDW* s = static_cast<SuspensionDW*>(c->suspensions[i]); PdDW& d = P->dw[i];
        v3(d.refPoint, s->dataRelToWheel.refPoint); v3(d.baseCarSteer, s->baseCarSteerPosition); v3(d.tyreSteer, s->dataRelToWheel.tyreSteer);
        d.rodLength = s->rodLength; d.k = s->k; d.progressiveK = s->progressiveK; d.packerRange = s->packerRange; d.bumpStopRate = s->bumpStopRate;
        d.bumpStopProgressive = s->bumpStopProgressive; d.bumpStopUp = s->bumpStopUp; d.bumpStopDn = s->bumpStopDn;
        d.toeOutLinear = s->toeOUT_Linear; d.staticCamber = s->staticCamber; d.baseCFM = s->baseCFM;
        copy_damper(d.damper, s->damper);
        for (int l = 0; l < PD_DW_LINKS; ++l) copy_dball(d.link[l], s->joints[l].get(), pdref_joint(s->joints[l].get())->targetDistance);
        { oder::Body* b = pdref_body(s->hub.get()); d.hubMass = b->mass; for (int k = 0; k < 3; ++k) d.hubInertia[k] = b->I[k]; }
        if (s->useActiveActuator) fprintf(stderr, "[oracle] DWB active actuator present: not exported\n");
    }
    for (int i = 0; i < 2 && !front_dw(c); ++i) {
        SuspensionStrut* s = static_cast<SuspensionStrut*>(c->suspensions[i]); PdStrut& d = P->strut[i];
        v3(d.refPoint, s->dataRelToWheel.refPoint); v3(d.carStrut, s->dataRelToBody.carStrut); v3(d.tyreStrut, s->dataRelToWheel.tyreStrut);
        v3(d.baseCarSteer, s->baseCarSteerPosition); v3(d.tyreSteer, s->dataRelToWheel.tyreSteer);
        d.strutBaseLength = s->strutBaseLength; d.strutBodyLength = s->strutBodyLength;
        d.rodLength = s->rodLength; d.k = s->k; d.progressiveK = s->progressiveK; d.packerRange = s->packerRange; d.bumpStopRate = s->bumpStopRate;
        d.bumpStopUp = s->bumpStopUp; d.bumpStopDn = s->bumpStopDn; d.toeOutLinear = s->toeOUT_Linear; d.staticCamber = s->staticCamber; d.baseCFM = s->baseCFM;
        copy_damper(d.damper, s->damper);
        for (int l = 0; l < 3; ++l) copy_dball(d.link[l], s->joints[l].get(), pdref_joint(s->joints[l].get())->targetDistance);
        { oder::Joint* j = pdref_joint(s->joints[3].get()); for (int k = 0; k < 3; ++k) { d.sliderAxis1[k] = j->axis1[k]; d.sliderOffset[k] = j->offset[k]; } for (int k = 0; k < 4; ++k) d.sliderQrel[k] = j->qrel[k]; }
        { oder::Joint* j = pdref_joint(s->joints[4].get()); for (int k = 0; k < 3; ++k) { d.ballAnchor1[k] = j->anchor1[k]; d.ballAnchor2[k] = j->anchor2[k]; } }
        { oder::Body* b = pdref_body(s->hub.get()); d.hubMass = b->mass; for (int k = 0; k < 3; ++k) d.hubInertia[k] = b->I[k]; }
        { oder::Body* b = pdref_body(s->strutBody.get()); d.strutMass = b->mass; for (int k = 0; k < 3; ++k) d.strutInertia[k] = b->I[k]; }
    }
    if (!rear_dw(c)) {
        SuspensionAxle* s = static_cast<SuspensionAxle*>(c->suspensions[2]); PdAxle& d = P->axle;
        d.track = s->track; d.referenceY = s->referenceY; d.attachRelativePos = s->attachRelativePos; v3(d.axleBasePos, s->axleBasePos); d.leafSpringKx = s->leafSpringK.x;
        d.rodLength = s->rodLength; d.k = s->k; d.progressiveK = s->progressiveK; d.bumpStopUp = s->bumpStopUp; d.bumpStopDn = s->bumpStopDn; d.bumpStopRate = s->bumpStopRate; d.baseCFM = s->baseCFM;
        copy_damper(d.damper, s->damper);
        d.nLinks = (int)s->joints.size();
        for (int l = 0; l < d.nLinks && l < PD_AXLE_LINKS; ++l) copy_dball(d.link[l], s->joints[l].ballAxle.joint.get(), pdref_joint(s->joints[l].ballAxle.joint.get())->targetDistance);
        { oder::Body* b = pdref_body(c->rigidAxle.get()); d.axleMass = b->mass; for (int k = 0; k < 3; ++k) d.axleInertia[k] = b->I[k]; }
        d.torqueReaction = c->axleTorqueReaction;
    }
    for (auto& hs : c->heaveSprings) {
        if (!hs->isPresent) continue;
        PdHeave& d = P->heave[hs->isFront ? 0 : 1];
        d.present = 1; d.bumpStopUp = hs->bumpStopUp; d.bumpStopDn = hs->bumpStopDn; d.rodLength = hs->rodLength; d.k = hs->k; d.progressiveK = hs->progressiveK;
        d.bumpStopRate = hs->bumpStopRate; d.packerRange = hs->packerRange; copy_damper(d.damper, hs->damper);
    }
    P->arbK[0] = c->antirollBars[0]->k; P->arbK[1] = c->antirollBars[1]->k;
    { BrakeSystem* b = c->brakeSystem.get(); P->brakes.brakePower = b->brakePower; P->brakes.brakePowerMultiplier = b->brakePowerMultiplier;
      P->brakes.handBrakeTorque = b->handBrakeTorque; P->brakes.frontBias = b->frontBias; P->brakes.biasMin = b->biasMin; P->brakes.biasMax = b->biasMax;
      P->brakes.ebbInternal = b->ebbMode == EBBMode::Internal ? 1 : 0; P->brakes.ebbFrontMultiplier = P->brakes.ebbInternal ? b->ebbFrontMultiplier : 0.0f;
      P->brakes.hasTemps = b->hasBrakeTempsData ? 1 : 0;
      if (b->hasBrakeTempsData) for (int i = 0; i < 4; ++i) { copy_curve(P->brakes.disc[i].perfCurve, b->discs[i].perfCurve); P->brakes.disc[i].torqueK = b->discs[i].torqueK; P->brakes.disc[i].coolTransfer = b->discs[i].coolTransfer; P->brakes.disc[i].coolSpeedFactor = b->discs[i].coolSpeedFactor; }
      if (b->ebbMode == EBBMode::DynamicController || b->steerBrake.isActive) fprintf(stderr, "[oracle] brake controllers present: not exported\n"); }
    for (int w = 0; w < 4; ++w) {
        Tyre* t = c->tyres[w].get(); PdTyre& d = P->tyre[w]; SCTM* m = t->tyreModel.get();
        d.width = t->data.width; d.radius = t->data.radius; d.rimRadius = t->data.rimRadius; d.k = t->data.k; d.d = t->data.d; d.angularInertia = t->data.angularInertia;
        d.thermalFrictionK = t->data.thermalFrictionK; d.thermalRollingK = t->data.thermalRollingK; d.thermalRollingSurfaceK = t->data.thermalRollingSurfaceK;
        d.softnessIndex = t->data.softnessIndex; d.radiusRaiseK = t->data.radiusRaiseK;
        d.grainThreshold = t->data.grainThreshold; d.blisterThreshold = t->data.blisterThreshold; d.grainGamma = t->data.grainGamma; d.blisterGamma = t->data.blisterGamma;
        d.grainGain = t->data.grainGain; d.blisterGain = t->data.blisterGain; d.optimumTemp = t->data.optimumTemp;
        d.version = t->modelData.version; d.Fz0 = t->modelData.Fz0; d.relaxationLength = t->modelData.relaxationLength;
        d.rr0 = t->modelData.rr0; d.rr1 = t->modelData.rr1; d.rr_slip = t->modelData.rr_slip;
        d.pressureSpringGain = t->modelData.pressureSpringGain; d.pressureRRGain = t->modelData.pressureRRGain; d.pressureGainD = t->modelData.pressureGainD;
        d.idealPressure = t->modelData.idealPressure; d.pressureRef = t->modelData.pressureRef;
        d.Dx0 = t->modelData.Dx0; d.Dx1 = t->modelData.Dx1; d.lsMultX = t->modelData.lsMultX; d.lsExpX = t->modelData.lsExpX;
        copy_curve(d.wearCurve, t->modelData.wearCurve);
        d.lsMultY = m->lsMultY; d.lsExpY = m->lsExpY; d.sctmLsMultX = m->lsMultX; d.sctmLsExpX = m->lsExpX; d.sctmFz0 = m->Fz0;
        d.maxSlip0 = m->maxSlip0; d.maxSlip1 = m->maxSlip1; d.asy = m->asy; d.falloffSpeed = m->falloffSpeed;
        d.speedSensitivity = m->speedSensitivity; d.camberGain = m->camberGain; d.dcamber0 = m->dcamber0; d.dcamber1 = m->dcamber1;
        d.cfXmult = m->cfXmult; d.pressureCfGain = m->pressureCfGain; d.brakeDXMod = m->brakeDXMod; d.dCamberBlend = m->dCamberBlend; d.combinedFactor = m->combinedFactor;
        if (m->dyLoadCurve.getCount() || m->dxLoadCurve.getCount() || m->dCamberCurve.getCount()) fprintf(stderr, "[oracle] tyre load/camber curves present: not exported\n");
        TyreThermalModel* th = t->thermalModel.get();
        d.surfaceTransfer = th->patchData.surfaceTransfer; d.patchTransfer = th->patchData.patchTransfer; d.patchCoreTransfer = th->patchData.patchCoreTransfer;
        d.internalCoreTransfer = th->patchData.internalCoreTransfer; d.coolFactorGain = th->patchData.coolFactorGain; d.camberSpreadK = th->camberSpreadK;
        copy_curve(d.performanceCurve, th->performanceCurve);
        d.flatSpotK = t->flatSpotK; d.explosionTemperature = t->explosionTemperature; d.pressureTemperatureGain = t->pressureTemperatureGain;
        d.pressureStaticDefault = t->status.pressureStatic;   /* compound value, possibly re-tuned (SetupManager PRESSURE_xx) */
        d.driven = t->driven ? 1 : 0; d.useLoadForVKM = t->useLoadForVKM ? 1 : 0;
    }
    P->nWings = (int)c->aeroMap->wings.size();
    for (int i = 0; i < P->nWings && i < PD_MAX_WINGS; ++i) {
        Wing* w = c->aeroMap->wings[i].get(); PdWing& d = P->wing[i];
        v3(d.position, w->data.position); d.area = w->data.area; d.cdGain = w->data.cdGain; d.clGain = w->data.clGain;
        d.angle = w->status.angle; d.angleMult = w->status.angleMult; d.yawGain = w->data.yawGain; d.isVertical = w->data.isVertical ? 1 : 0;
        copy_curve(d.lutAOA_CL, w->data.lutAOA_CL); copy_curve(d.lutAOA_CD, w->data.lutAOA_CD);
        if (w->data.lutGH_CL.getCount() || w->data.lutGH_CD.getCount() || !w->dynamicControllers.empty()) fprintf(stderr, "[oracle] wing GH luts / controllers present: not exported\n");
    }
    {
        Engine* e = c->drivetrain->engineModel.get(); PdEngine& d = P->engine;
        copy_curve(d.powerCurve, e->data.powerCurve); copy_curve(d.throttleResponseCurve, e->throttleResponseCurve);
        d.minimum = e->data.minimum; d.limiter = e->data.limiter; d.limiterCycles = e->data.limiterCycles;
        d.coast1 = e->data.coast1; d.coast2 = e->data.coast2; d.inertia = e->inertia; d.limiterMultiplier = e->limiterMultiplier;
        d.rpmDamageThreshold = e->rpmDamageThreshold; d.rpmDamageK = e->rpmDamageK; d.turboBoostDamageThreshold = e->turboBoostDamageThreshold;
        d.turboBoostDamageK = e->turboBoostDamageK; d.bovThreshold = e->bovThreshold;
        d.gasCoastOffset = e->gasCoastOffset; d.coastEntryRpm = e->coastEntryRpm;
        d.overlapFreq = e->data.overlapFreq; d.overlapGain = e->data.overlapGain; d.overlapIdealRPM = e->data.overlapIdealRPM;
        d.isEngineStallEnabled = e->isEngineStallEnabled ? 1 : 0; d.maxPowerRPM = e->maxPowerRPM; d.maxTorqueRPM = e->maxTorqueRPM;
        P->nTurbos = (int)e->turbos.size();
        for (int i = 0; i < P->nTurbos && i < PD_MAX_TURBOS; ++i) {
            const Turbo& t = *e->turbos[i]; PdTurbo& u = P->turbo[i];
            u.lagDN = t.data.lagDN; u.lagUP = t.data.lagUP; u.maxBoost = t.data.maxBoost; u.wastegate = t.data.wastegate; u.rpmRef = t.data.rpmRef; u.gamma = t.data.gamma;
            u.userSetting = t.userSetting; u.isAdjustable = t.data.isAdjustable ? 1 : 0;
        }
        copy_curve(d.throttleResponseCurveMax, e->throttleResponseCurveMax); d.throttleResponseCurveMaxRef = e->throttleResponseCurveMax.getCount() ? e->throttleResponseCurveMaxRef : 0.0f;
        if (!e->turboControllers.empty()) fprintf(stderr, "[oracle] turbo controllers present: not exported\n");
    }
    {
        Drivetrain* t = c->drivetrain.get(); PdDrivetrain& d = P->drivetrain;
        d.nGears = (int)t->gears.size(); for (int i = 0; i < d.nGears && i < PD_MAX_GEARS; ++i) d.gears[i] = t->gears[i].ratio;
        d.tractionType = (int)t->tractionType; d.diffType = (int)t->diffType;
        d.finalRatio = t->finalRatio; d.diffPowerRamp = t->diffPowerRamp; d.diffCoastRamp = t->diffCoastRamp; d.diffPreLoad = t->diffPreLoad;
        d.gearUpTime = t->gearUpTime; d.gearDnTime = t->gearDnTime; d.autoCutOffTime = t->autoCutOffTime; d.controlsWindowGain = t->controlsWindowGain;
        d.orgRpmWindow = t->orgRpmWindow; d.damageRpmWindow = t->damageRpmWindow; d.clutchMaxTorque = t->clutchMaxTorque; d.clutchInertia = t->clutchInertia;
        d.driveInertia = t->drive.inertia; d.shaftInertiaL = t->outShaftL.inertia; d.shaftInertiaR = t->outShaftR.inertia;
        d.isShifterSupported = t->isShifterSupported ? 1 : 0;
    }
    {
        PdAssists& d = P->assists; AutoClutch* ac = c->autoClutch.get(); AutoBlip* ab = c->autoBlip.get(); AutoShifter* as = c->autoShift.get();
        copy_curve(d.upshiftProfile, ac->upshiftProfile); copy_curve(d.downshiftProfile, ac->downshiftProfile); copy_curve(d.blipProfile, ab->blipProfile);
        d.acRpmMin = ac->rpmMin; d.acRpmMax = ac->rpmMax; d.acClutchSpeed = ac->clutchSpeed;
        d.acUseAutoOnStart = ac->useAutoOnStart ? 1 : 0; d.acUseAutoOnChange = ac->useAutoOnChange ? 1 : 0; d.acIsForced = ac->isForced ? 1 : 0;
        d.blipPerformTime = ab->blipPerformTime; d.blipIsActive = ab->isActive ? 1 : 0; d.blipIsElectronic = ab->isElectronic ? 1 : 0;
        d.asChangeUpRpm = as->changeUpRpm; d.asChangeDnRpm = as->changeDnRpm; d.asSlipThreshold = as->slipThreshold; d.asGasCutoffTime = as->gasCutoffTime; d.asIsActive = as->isActive ? 1 : 0;
    }
    P->nProbes = (int)c->probes.size();
    for (int i = 0; i < P->nProbes && i < PD_MAX_PROBES; ++i) { v3(P->probeDir[i], c->probes[i].dir); P->probeLength[i] = c->probes[i].length; }
    P->lookAheadCount = c->lookAheadCount; P->lookAheadStep = c->lookAheadStep;
    { auto* cfg = c->scoring->config; for (int i = 0; i < PD_NUM_SCORING_VARS && i < (int)cfg->vvars.size(); ++i) P->scoring[i] = cfg->vvars[i]->value; }
    P->teleportOnCollision = c->teleportOnCollision; P->teleportOnBadLocation = c->teleportOnBadLocation; P->teleportMode = c->teleportMode;
    P->ambientTemperature = sim->ambientTemperature; P->roadTemperature = sim->roadTemperature; P->airDensity = sim->getAirDensity();
    P->fuelConsumptionRate = sim->fuelConsumptionRate; P->tyreConsumptionRate = sim->tyreConsumptionRate; P->mechanicalDamageRate = sim->mechanicalDamageRate;
    P->allowTyreBlankets = sim->allowTyreBlankets ? 1 : 0;
    { oder::World* w = pdref_world(sim->physics.get()); P->gravityY = w->gravity[1]; P->worldERP = w->erp; P->worldCFM = w->cfm; }
    { /* colliders exactly as the physics engine received them (CarColliderManager.cpp:32, Car.cpp:377) */
        P->hasBoxCollider = pdref_body_box(c->body.get(), P->boxCentre, P->boxSize) > 0 ? 1 : 0;
        ITriMesh* mesh = nullptr; float off[3] = {0, 0, 0}, r[9];
        if (pdref_body_mesh(c->body.get(), &mesh, r, off) > 0 && mesh) {
            P->nColliderVerts = (int)mesh->getVertexCount(); P->nColliderTris = (int)(mesh->getIndexCount() / 3);
            for (int k = 0; k < 3; ++k) { P->colliderMin[k] = 3.4e38f; P->colliderMax[k] = -3.4e38f; }
            const TriMeshVertex* vb = mesh->getVB(); const TriMeshIndex* ib = mesh->getIB();
            for (int i = 0; i < P->nColliderVerts && i < PD_MAX_COLLIDER_VERTS; ++i) {
                const float v[3] = {(r[0] * vb[i].x + r[1] * vb[i].y + r[2] * vb[i].z) + off[0], (r[3] * vb[i].x + r[4] * vb[i].y + r[5] * vb[i].z) + off[1], (r[6] * vb[i].x + r[7] * vb[i].y + r[8] * vb[i].z) + off[2]};
                for (int k = 0; k < 3; ++k) { P->colliderVerts[i][k] = v[k]; P->colliderMin[k] = std::min(P->colliderMin[k], v[k]); P->colliderMax[k] = std::max(P->colliderMax[k], v[k]); }
            }
            for (int t = 0; t < P->nColliderTris && t < PD_MAX_COLLIDER_TRIS; ++t) for (int k = 0; k < 3; ++k) P->colliderTris[t][k] = (uint8_t)ib[t * 3 + k];
            for (int t = 0; t < P->nColliderTris && t < PD_MAX_COLLIDER_TRIS; ++t) for (int k = 0; k < 3; ++k) {   /* derived filter data of the product's layout */
                const float a = P->colliderVerts[P->colliderTris[t][0]][k], b = P->colliderVerts[P->colliderTris[t][1]][k], c2 = P->colliderVerts[P->colliderTris[t][2]][k];
                P->colliderTriBounds[t][k] = std::min(a, std::min(b, c2)) - 1e-4f; P->colliderTriBounds[t][3 + k] = std::max(a, std::max(b, c2)) + 1e-4f;
            }
            for (int t = 0; t < P->nColliderTris && t < PD_MAX_COLLIDER_TRIS; ++t) {
                float cc[3], r2 = 0.0f;
                for (int k = 0; k < 3; ++k) cc[k] = 0.5f * (P->colliderTriBounds[t][k] + P->colliderTriBounds[t][3 + k]);
                for (int v = 0; v < 3; ++v) { float d2 = 0.0f; for (int k = 0; k < 3; ++k) { const float d = P->colliderVerts[P->colliderTris[t][v]][k] - cc[k]; d2 += d * d; } r2 = std::max(r2, d2); }
                P->colliderTriSphere[t][0] = cc[0]; P->colliderTriSphere[t][1] = cc[1]; P->colliderTriSphere[t][2] = cc[2]; P->colliderTriSphere[t][3] = sqrtf(r2) * 1.0001f + 1e-4f;
            }
        }
    }
}

/* track-level derived data (Track::initTrackPoints, Track.cpp:178-272) for loader parity */
void pdref_get_track_info(void* hv, PdTrackInfo* T) {
    RefSim* h = (RefSim*)hv; Track* t = h->sim->track.get();
    memset(T, 0, sizeof(*T));
    T->nSurfaces = (int)t->surfaces.size();
    int nt = 0; for (auto& s : t->surfaces) nt += (int)(s->trimesh->getIndexCount() / 3);
    T->nTris = nt; T->nFatPoints = (int)t->fatPoints.size();
    T->nSplineNodes = t->interpolatedSpline ? t->interpolatedSpline->node_count() : 0;
    T->interpolateStep = t->interpolateStep; T->closedLoop = t->closedLoop ? 1 : 0;
    T->computedTrackLength = t->computedTrackLength; T->computedTrackWidth = t->computedTrackWidth;
    T->dynamicGripLevel = t->dynamicGripLevel; T->hashCellSize = t->fatPointsHash.cellSize;
}
void pdref_get_spline_nodes(void* hv, float* xyz, float* dist) {
    Track* t = ((RefSim*)hv)->sim->track.get();
    for (int i = 0; i < t->interpolatedSpline->node_count(); ++i) {
        const vec3f& n = t->interpolatedSpline->node(i); xyz[i * 3] = n.x; xyz[i * 3 + 1] = n.y; xyz[i * 3 + 2] = n.z;
        dist[i] = t->interpolatedSpline->length_at_point(i);
    }
}

/* ---- component hooks (unit parity) ---- */
/* in: n x 9 {load, slipAngleRAD, slipRatio, camberRAD, speed, u, cpLength, grain, blister|pressureRatio packed separately}
 * layout used: [load, sa, sr, camber, speed, u, cpLength, pressureRatio, blister];  out: n x 7 TyreModelOutput */
void pdref_sctm_solve(void* hv, int wheel, int n, const float* in, float* out) {
    SCTM* m = ((RefSim*)hv)->car->tyres[wheel]->tyreModel.get();
    for (int i = 0; i < n; ++i) {
        TyreModelInput t; const float* p = in + i * 9;
        t.load = p[0]; t.slipAngleRAD = p[1]; t.slipRatio = p[2]; t.camberRAD = p[3]; t.speed = p[4]; t.u = p[5]; t.cpLength = p[6];
        t.pressureRatio = p[7]; t.blister = p[8]; t.grain = 0; t.tyreIndex = wheel; t.useSimpleModel = false;
        TyreModelOutput o = m->solve(t);
        float* q = out + i * 7; q[0] = o.Fy; q[1] = o.Fx; q[2] = o.Mz; q[3] = o.trail; q[4] = o.ndSlip; q[5] = o.Dy; q[6] = o.Dx;
    }
}
/* rays: n x 7 {ox,oy,oz,dx,dy,dz,len} -> n x 8 {hit, px,py,pz, nx,ny,nz, surfaceIndex} */
void pdref_raycast(void* hv, int n, const float* in, float* out) {
    RefSim* h = (RefSim*)hv; Track* t = h->sim->track.get();
    for (int i = 0; i < n; ++i) {
        const float* p = in + i * 7; float* q = out + i * 8;
        RayCastHit hit = h->sim->physics->rayCast(vec3f(p[0], p[1], p[2]), vec3f(p[3], p[4], p[5]), p[6]);
        q[0] = hit.hasContact ? 1.0f : 0.0f; q[1] = hit.pos.x; q[2] = hit.pos.y; q[3] = hit.pos.z; q[4] = hit.normal.x; q[5] = hit.normal.y; q[6] = hit.normal.z;
        q[7] = hit.hasContact ? (float)surface_index(t, (Surface*)hit.collisionObject->getUserPointer()) : -1.0f;
    }
}
/* ---- independent pins of the ODE restatement: quantities that do not depend on how the restatement is written ---- */
/* largest constraint violation |C(q)| over the car's joints after the last step: ball / fixed anchors apart (m), dball length error
 * (m), slider: offset across the axis (m); out3 = {max linear error, max dball error, number of joints} */
void pdref_constraint_errors(void* hv, double* out3) {
    oder::World* w = pdref_world(((RefSim*)hv)->sim->physics.get());
    double lin = 0, db = 0;
    for (oder::Joint* j : w->joints) {
        const oder::Body& b0 = *j->b0; const oder::Body& b1 = *j->b1;
        if (j->type == oder::J_BALL) {
            float g1[3], g2[3]; oder::body_rel_point_pos(b0, j->anchor1, g1); oder::body_rel_point_pos(b1, j->anchor2, g2);
            const double e = sqrt((double)(g1[0] - g2[0]) * (g1[0] - g2[0]) + (double)(g1[1] - g2[1]) * (g1[1] - g2[1]) + (double)(g1[2] - g2[2]) * (g1[2] - g2[2]));
            if (e > lin) lin = e;
        } else if (j->type == oder::J_DBALL) {
            const double e = fabs((double)oder::dball_current_distance(*j) - (double)j->targetDistance);
            if (e > db) db = e;
        } else if (j->type == oder::J_FIXED) {
            float ofs[3]; oder::mul0_331(ofs, b0.R, j->offset);
            const double ex = b1.pos[0] - b0.pos[0] + ofs[0], ey = b1.pos[1] - b0.pos[1] + ofs[1], ez = b1.pos[2] - b0.pos[2] + ofs[2];
            const double e = sqrt(ex * ex + ey * ey + ez * ez);
            if (e > lin) lin = e;
        } else if (j->type == oder::J_SLIDER) {
            float ax1[3], p[3], q[3], ofs[3]; oder::mul0_331(ax1, b0.R, j->axis1); oder::plane_space(ax1, p, q); oder::mul0_331(ofs, b1.R, j->offset);
            const float c[3] = {b1.pos[0] - b0.pos[0] + ofs[0], b1.pos[1] - b0.pos[1] + ofs[1], b1.pos[2] - b0.pos[2] + ofs[2]};
            const double e = sqrt((double)oder::dot3(p, c) * oder::dot3(p, c) + (double)oder::dot3(q, c) * oder::dot3(q, c));
            if (e > lin) lin = e;
        }
    }
    out3[0] = lin; out3[1] = db; out3[2] = (double)w->joints.size();
}
/* the last step's bilateral system (A before factoring, rhs) solved again in DOUBLE precision by Gaussian elimination with partial
 * pivoting, against the single-precision lambda the restated LDL^T produced: out2 = {max |lambda32 - lambda64| / max |lambda64|, rows} */
void pdref_keep_system(void* hv, int on) { pdref_world(((RefSim*)hv)->sim->physics.get())->keepSystem = on != 0; }
void pdref_resolve_fp64(void* hv, double* out2) {
    oder::World* w = pdref_world(((RefSim*)hv)->sim->physics.get());
    const int m = (int)w->lastRhs.size();
    out2[0] = -1; out2[1] = m;
    if (m == 0 || (int)w->last_lambda.size() < m) return;
    std::vector<double> A((size_t)m * m), x(m);
    for (size_t k = 0; k < A.size(); ++k) A[k] = w->lastA[k];
    for (int i = 0; i < m; ++i) x[i] = w->lastRhs[i];
    for (int c = 0; c < m; ++c) {
        int piv = c; for (int r = c + 1; r < m; ++r) if (fabs(A[(size_t)r * m + c]) > fabs(A[(size_t)piv * m + c])) piv = r;
        if (piv != c) { for (int k = 0; k < m; ++k) std::swap(A[(size_t)c * m + k], A[(size_t)piv * m + k]); std::swap(x[c], x[piv]); }
        const double d = A[(size_t)c * m + c];
        for (int r = c + 1; r < m; ++r) { const double f = A[(size_t)r * m + c] / d; if (f == 0) continue; for (int k = c; k < m; ++k) A[(size_t)r * m + k] -= f * A[(size_t)c * m + k]; x[r] -= f * x[c]; }
    }
    for (int r = m - 1; r >= 0; --r) { double sacc = x[r]; for (int k = r + 1; k < m; ++k) sacc -= A[(size_t)r * m + k] * x[k]; x[r] = sacc / A[(size_t)r * m + r]; }
    double num = 0, den = 0;
    for (int i = 0; i < m; ++i) { num = std::max(num, fabs((double)w->last_lambda[i] - x[i])); den = std::max(den, fabs(x[i])); }
    out2[0] = den > 0 ? num / den : 0;
}
/* a free rigid body (no joints, no gravity) tumbling about a non-principal axis, stepped by the restated dxStepIsland / dxStepBody:
 * relative drift of the angular momentum's magnitude, of the kinetic energy and of the linear momentum after `steps` steps */
void pdref_free_body_drift(int steps, double h, double* out3) {
    oder::World w; w.gravity[0] = w.gravity[1] = w.gravity[2] = 0;
    oder::Body b; oder::body_set_mass_box(b, 837.4f, 1.40f, 1.35f, 4.70f);          /* the demo car's chassis (car.ini [BASIC]) */
    b.avel[0] = 0.7f; b.avel[1] = 1.1f; b.avel[2] = -0.4f; b.lvel[0] = 3.0f; b.lvel[1] = 0.5f; b.lvel[2] = -12.0f;
    w.bodies.push_back(&b);
    auto momentum = [&](double* L, double& E) {
        float Iw[9], inv[9]; oder::world_inertia(b, Iw, inv);
        for (int i = 0; i < 3; ++i) L[i] = (double)Iw[i * 3] * b.avel[0] + (double)Iw[i * 3 + 1] * b.avel[1] + (double)Iw[i * 3 + 2] * b.avel[2];
        E = 0.5 * (L[0] * b.avel[0] + L[1] * b.avel[1] + L[2] * b.avel[2]);
    };
    double L0[3], E0; momentum(L0, E0);
    const double p0[3] = {b.lvel[0], b.lvel[1], b.lvel[2]};
    for (int s = 0; s < steps; ++s) oder::world_step(w, (float)h);
    double L1[3], E1; momentum(L1, E1);
    const double n0 = sqrt(L0[0] * L0[0] + L0[1] * L0[1] + L0[2] * L0[2]), n1 = sqrt(L1[0] * L1[0] + L1[1] * L1[1] + L1[2] * L1[2]);
    out3[0] = fabs(n1 - n0) / n0; out3[1] = fabs(E1 - E0) / E0;
    out3[2] = sqrt((b.lvel[0] - p0[0]) * (b.lvel[0] - p0[0]) + (b.lvel[1] - p0[1]) * (b.lvel[1] - p0[1]) + (b.lvel[2] - p0[2]) * (b.lvel[2] - p0[2]));
}

/* The bench workload (bench.py: configs[1] sample) for `nsims` simulators, entirely on this side of the FFI, so that the CPU
 * figure carries no Python / ctypes time: random actions resampled every 33 ticks (xorshift, seeded), setCarControls +
 * stepSimulator per tick, env-style auto reset on collisionFlag / outOfTrackFlag (teleportToSpline(u), u random).  `preroll`
 * untimed ticks first.  Returns the seconds spent in the `ticks` timed ticks. */
double pdref_bench_loop(void** hv, int nsims, int ticks, int preroll, unsigned int seed, long long* resets_out) {
    unsigned long long st = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)seed * 0xD1B54A32D192ED03ull);
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (float)((st >> 40) & 0xFFFFFF) * (1.0f / 16777216.0f); };
    std::vector<float> steer((size_t)nsims, 0.0f), gas((size_t)nsims, 0.5f);
    long long resets = 0;
    struct timespec t0, t1; clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int t = -preroll; t < ticks; ++t) {
        if (t == 0) clock_gettime(CLOCK_MONOTONIC, &t0);
        if ((t + preroll) % 33 == 0) for (int i = 0; i < nsims; ++i) { steer[(size_t)i] = rnd() * 2.0f - 1.0f; gas[(size_t)i] = 0.1f + 0.9f * rnd(); }
        for (int i = 0; i < nsims; ++i) {
            RefSim* h = (RefSim*)hv[i]; Car* c = h->car;
            CarControls k; k.steer = steer[(size_t)i]; k.gas = gas[(size_t)i]; k.isShifterSupported = c->controls.isShifterSupported;
            c->controls = k; c->smoothSteer = true;
            Simulator* s = h->sim.get();
            s->step(1.0f / 333.0f, s->physicsTime, s->gameTime);
            s->physicsTime += 1.0 / 333.0; s->gameTime += 1.0 / 333.0;
            if (c->collisionFlag || c->outOfTrackFlag) { c->teleportToSpline(rnd()); ++resets; }
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (resets_out) *resets_out = resets;
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
int pdref_contacts(void* hv, float* out8, int cap) { return pdref_engine_contacts(((RefSim*)hv)->sim->physics.get(), out8, cap); }
void pdref_set_collision_response(void* hv, int on) { pdref_engine_set_response(((RefSim*)hv)->sim->physics.get(), on); }
unsigned int pdref_frame(void* hv) { return pdref_get_frame(((RefSim*)hv)->sim->physics.get()); }
void pdref_set_frame_counter(void* hv, unsigned int f) { pdref_set_frame(((RefSim*)hv)->sim->physics.get(), f); }
void pdref_ray_statistics(void* hv, unsigned long long* rays, unsigned long long* tris) { pdref_ray_stats(((RefSim*)hv)->sim->physics.get(), rays, tris); }
int pdref_num_rows(void* hv) { return pdref_world(((RefSim*)hv)->sim->physics.get())->last_m; }

} /* extern "C" */
