/*
 * pd_batch.h -- C ABI of the B200 batched vehicle-physics core (libpd_b200.so).
 *
 * This is the drop-in boundary for the Car::step hot path: plain pointers and sizes, no C++ or torch types.
 * One `pd_batch` plays the role of N reference simulators with one car each (the way
 * pyprojectd/projectd_env.py:118-121 uses the reference: one Simulator + one Car per environment), all
 * sharing one track and one car model, advanced together by CUDA kernels on one GPU.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference tree).
 * Every function returns 0 on success and a negative code on failure; nothing throws across the
 * boundary (the reference's binding catches std::exception and returns -1 / no-ops:
 * src/PyProjectD/PyProjectD.cpp:132-136,198-201,232-236).  pd_last_error() gives the message.
 * There is NO CPU fallback: if no CUDA device is usable pd_create fails.
 *
 * Pointers marked [host|device] are interpreted according to the `on_device` argument of the call.
 */
#ifndef PD_BATCH_H
#define PD_BATCH_H

#include <stdint.h>
#include "pd_state.h"
#include "pd_params.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pd_batch pd_batch;

#define PD_OK            0
#define PD_ERR_ARG      -1
#define PD_ERR_IO       -2   /* missing / malformed car or track data */
#define PD_ERR_CUDA     -3
#define PD_ERR_UNSUPPORTED -4

/* teleport modes: Car/Car.h:19-24 (TeleportMode) */
#define PD_TELEPORT_START   0
#define PD_TELEPORT_NEAREST 1
#define PD_TELEPORT_RANDOM  2

/* termination causes reported by pd_env_step (pyprojectd/projectd_env.py:186-206) */
#define PD_DONE_COLLISION  1
#define PD_DONE_OFFTRACK   2
#define PD_DONE_STUCK      4
#define PD_DONE_LOWREWARD  8
#define PD_DONE_NAN       16

/* when pd_env_step resets a finished env (see pd_set_autoreset) */
#define PD_AUTORESET_SAME_STEP 0   /* inside the step that finished it (what SB3's VecEnv does around ProjectDEnv): 3 launches */
#define PD_AUTORESET_NEXT_STEP 1   /* at the following step, whose action is ignored (gymnasium >= 1.0 VectorEnv): 1 launch */

/* createSimulator + loadTrack + addCar for n_envs environments
 * (PyProjectD.cpp:111-137 createSimulator, :186-203 loadTrack, :219-237 addCar; Simulator::init
 * Sim/Simulator.cpp:22-90, Track::init Sim/Track.cpp:27-50, Car::init Car/Car.cpp:31-223).
 * base_path is the reference's base directory (holds cfg/ and content/).  device = CUDA ordinal. */
int pd_create(const char* base_path, const char* track_name, const char* car_model, int n_envs, int device, pd_batch** out);
/* same, but the track is a generated closed circuit of about `target_tris` triangles (BASELINE.json config 4) */
int pd_create_synthetic(const char* base_path, const char* car_model, int target_tris, float length_m, int n_envs, int device, pd_batch** out);
/* Track::computeFatPoints + computeSideLocation (Sim/Track.cpp:366-467) as batch ray casting on the GPU: regenerates the content of
 * spline.cache (15 floats per spline point: best, left, right, center, forwardDir) from spline.bin + surfaces.bin, ignoring any
 * cache on disk.  Returns the number of points (out holds min(points, cap_points)) or a negative PD_ERR_*.  pd_create does the same
 * by itself when a track has no usable spline.cache. */
int pd_compute_fat_points(const char* base_path, const char* track_name, int device, float* out, int cap_points);
/* destroySimulator (PyProjectD.cpp:139-149) */
void pd_destroy(pd_batch* b);
const char* pd_last_error(const pd_batch* b);      /* b may be NULL: error of the last failed pd_create */

int pd_num_envs(const pd_batch* b);
int pd_state_words(void);                           /* PD_STATE_WORDS of this build */
int pd_obs_dim(void);                               /* 24 */
int pd_car_state_bytes(void);                       /* sizeof(CarState) = 664 (Car/CarState.h:11-56) */

/* setCarAssists (PyProjectD.cpp:307-317) -- applies to every env */
int pd_set_assists(pd_batch* b, int auto_clutch, int auto_shift, int auto_blip);
/* setCarTune (PyProjectD.cpp:328-335 -> SetupManager::setTune, Car/SetupManager.cpp:283-288) */
int pd_set_tune(pd_batch* b, const char* name, float value);
/* setCarRawTune (PyProjectD.cpp:337-344 -> SetupManager::setRawTune, :276-281): the raw value, no spinner clamp / multiplier.
 * Both return PD_ERR_UNSUPPORTED for a variable the reference registers but this build cannot honour (rear per-wheel suspension
 * tunes: the rigid axle keeps one parameter set; AWD differentials; turbos); names the reference does not know are ignored as there. */
int pd_set_raw_tune(pd_batch* b, const char* name, float value);
/* setScoringVar / getScoringVar (PyProjectD.cpp:346-363; process-global ScoringConfig in the reference) */
int pd_set_scoring_var(pd_batch* b, const char* name, float value);
float pd_get_scoring_var(const pd_batch* b, const char* name);

/* setCarControls (PyProjectD.cpp:297-305): controls[n_envs][5] = steer, clutch, brake, handBrake, gas;
 * gears[n_envs][3] = requestedGearIndex (-1 = sequential), gearUp, gearDn, or NULL for (-1, 0, 0).
 * `smooth` = the `smooth` flag of setCarControls. */
int pd_set_controls(pd_batch* b, const float* controls, const int8_t* gears, int smooth, int on_device);
/* the env's action mapping (pyprojectd/projectd_env.py:159-160): actions[n_envs][2] in [-1,1] ->
 * steer = a0, gas = linscale(a1, -1, 1, 0.1, 1), other controls 0, sequential gearbox, smooth steering */
int pd_set_actions(pd_batch* b, const float* actions, int on_device);

/* stepSimulator (PyProjectD.cpp:160-180 -> Simulator::step, Sim/Simulator.cpp:168-201), n_ticks times:
 * Car::step + dWorldStep + Car::postStep for every env; physicsTime += dt after each tick. */
int pd_step(pd_batch* b, float dt, int n_ticks);
double pd_get_time(const pd_batch* b);
int pd_set_time(pd_batch* b, double t);

/* teleportCarToSpline / teleportCarByMode (PyProjectD.cpp:268-284 -> Car::teleportToSpline / teleportByMode,
 * Car/Car.cpp:1325-1358).  mask[n_envs] (NULL = all envs, host memory); dist_norm[n_envs] (host) or NULL.
 * PD_TELEPORT_RANDOM draws u ~ U[0,1) from a counter-based generator keyed by (seed, global env id) so that
 * results do not depend on how envs are sharded over GPUs (the reference uses the process-global rand()). */
int pd_teleport_spline(pd_batch* b, const uint8_t* mask, const float* dist_norm);
int pd_teleport_mode(pd_batch* b, const uint8_t* mask, int mode);   /* `mode` also becomes the mode of pd_env_step's automatic resets (ProjectDEnv.teleport_mode, projectd_env.py:39,219) */
int pd_set_seed(pd_batch* b, uint64_t seed, uint64_t env_id_offset);   /* setSeed (PyProjectD.cpp:50-53) */

/* getCarState (PyProjectD.cpp:319-326): fills one 664-byte CarState (Car/CarState.h:11-56 layout) */
int pd_get_car_state(pd_batch* b, int env, void* car_state_out);
/* observations of the env (pyprojectd/projectd_env.py:237-275): out[n_envs][24] float32 */
int pd_get_obs(pd_batch* b, float* out, int to_device);
/* the same buffer as a DLPack capsule payload (DLManagedTensor*, device kDLCUDA, shape [n_envs,24], f32).
 * The tensor aliases the batch's observation buffer and is refreshed by pd_observe / pd_env_step. */
void* pd_obs_dlpack(pd_batch* b);
int pd_observe(pd_batch* b);                         /* recompute the observation buffer from the state */
const float* pd_obs_device_ptr(pd_batch* b);
/* stepReward[n_envs] f32 and flags[n_envs] i32 (bit0 collisionFlag, bit1 outOfTrackFlag) */
int pd_get_rewards(pd_batch* b, float* step_reward, float* total_reward, int32_t* flags);

/* The env-level knobs of ProjectDEnv (pyprojectd/projectd_env.py:27-53) that pd_env_step / pd_set_actions apply inside the kernels.
 * Defaults = the reference's class attributes. */
typedef struct PdEnvConfig {
    float min_gas, max_gas;                       /* gas = linscale(a1, -1..1 -> min_gas..max_gas)                 (:52-53,160) */
    float terminate_hit_penalty, terminate_off_track_penalty, terminate_stuck_penalty;               /* (:42-44) */
    float terminate_low_reward, stuck_timeout;                                                      /* (:46-47) */
    int32_t terminate_on_hit, terminate_off_track, terminate_when_stuck;                            /* (:38-40) */
    int32_t smooth_controls;                      /* the `smooth` argument of setCarControls                       (:27,168) */
    float clutch;                                 /* controls.clutch sent with every action: 0, or 1.0 when auto_clutch is off (:162-163) */
    int32_t requested_gear;                       /* controls.requestedGearIndex: -1 sequential, or 2 when auto_shift is off  (:165-166) */
} PdEnvConfig;
int pd_set_env_config(pd_batch* b, const PdEnvConfig* cfg);
int pd_get_env_config(const pd_batch* b, PdEnvConfig* out);
/* ProjectDEnv.reset's bookkeeping (projectd_env.py:224-227): zero the running episode return / length of every env (mask NULL)
 * or of the masked envs, after the reset's own step */
int pd_env_reset_counters(pd_batch* b, const uint8_t* mask);

/* One vectorised ProjectDEnv.step (pyprojectd/projectd_env.py:157-212) for all envs: set actions, one tick,
 * observations, reward with the env's termination penalties, done flags, and automatic reset
 * (teleport by `PD_TELEPORT_*` mode + one zero-action tick, projectd_env.py:216-227) of finished envs.
 * actions / obs / reward / done are DEVICE pointers (obs may be NULL to use the internal buffer). */
int pd_env_step(pd_batch* b, const float* actions_dev, float dt, float* obs_dev, float* reward_dev, int32_t* done_dev);
/* Auto-reset convention of pd_env_step.  PD_AUTORESET_SAME_STEP (default): a finished env is teleported and
 * advanced by the reset's zero-action tick right after the step that finished it; obs then holds the reset
 * observation (the reference env driven through a same-step auto-resetting vector wrapper).
 * PD_AUTORESET_NEXT_STEP: the step that finishes an env returns its terminal observation with done != 0; the
 * NEXT pd_env_step ignores that env's action, performs the reset (teleport + zero-action tick) inside the same
 * kernel launch as everybody else's tick and returns the reset observation with reward 0 and done 0. */
int pd_set_autoreset(pd_batch* b, int mode);
/* The same step for a caller that lives on the host (what ProjectDEnv.step is to a Python user): actions[n_envs][2]
 * are copied host -> device, obs[n_envs][24] / reward[n_envs] / done[n_envs] device -> host, all on the batch's
 * stream, one synchronisation at the end.  Pinned (page-locked) host buffers make the copies asynchronous;
 * pageable ones work too.  obs / reward / done may be NULL. */
int pd_env_step_host(pd_batch* b, const float* actions_host, float dt, float* obs_host, float* reward_host, int32_t* done_host);
/* episode statistics accumulated by pd_env_step since the last call: sums over envs, ready for an NCCL
 * all-reduce: {episodes, sum_return, sum_length, collisions, offtrack, stuck, lowreward, nan} as float64[8] */
int pd_env_stats(pd_batch* b, double* out8, int reset);

/* full per-env state record (include/pd_state.h), for parity tests, checkpoints and "identical states in" */
int pd_get_state(pd_batch* b, int env, uint32_t* record);
int pd_set_state(pd_batch* b, int env, const uint32_t* record);
int pd_snapshot(pd_batch* b, uint32_t* host_buf);    /* PD_STATE_WORDS * n_envs words, SoA as on the device */
int pd_restore(pd_batch* b, const uint32_t* host_buf);
int pd_get_params(const pd_batch* b, PdCarParams* out);
int pd_get_track_info(const pd_batch* b, PdTrackInfo* out);

/* batch ray cast against the track BVH (IPhysicsEngine::rayCast, Physics/IPhysicsEngine.h:24):
 * rays[n][7] = origin, direction, length -> out[n][8] = hit, pos, normal, surface index.  Host pointers. */
int pd_raycast(pd_batch* b, int n, const float* rays, float* out);

/* profiling aid: SM cycles every warp spent in the last full tick launch (batch created with env PD_DEBUG_CLOCKS=1);
 * returns the number of entries written (0 when disabled) */
int pd_debug_read_clocks(pd_batch* b, long long* out, int cap);

/* Collision RESPONSE (PhysicsEngineODE::onCollision, Physics/ODE/PhysicsEngineODE.cpp:284-341): when on (default),
 * a car that touches the static world on an odd physics frame gets contact joints (normal row + friction pyramid) that act in that
 * frame's solve and the next one, as the reference's two-frame contact groups do.  Off: contacts only raise collisionFlag (what an
 * env that terminates on hit needs; the detection then runs inside the tick kernel and costs no extra launch). */
int pd_set_collision_response(pd_batch* b, int on);
/* live contact joints of one env: out[i][8] = position, normal, depth, kind (0 floor box vs TRACK, 1 hull vs WALL); returns their number */
int pd_get_contacts(pd_batch* b, int env, float* out, int max_contacts);
/* sizeof(PdCarParams) of this build (size the buffer of pd_get_params from it) */
int pd_params_bytes(void);
/* Make every later kernel / copy of this batch run on `stream` (a cudaStream_t of the batch's device; NULL = back to the batch's own
 * non-blocking stream).  The previous stream is synchronised first.  With the caller's stream the batch's work is ordered with the
 * caller's own kernels (e.g. the torch ops that produce the actions and consume the observations) without events. */
int pd_set_stream(pd_batch* b, void* stream);
int pd_sync(pd_batch* b);
void* pd_stream(pd_batch* b);                        /* the cudaStream_t every kernel of this batch runs on */
/* number of kernels launched by this batch since creation (bench.py's gpu_launches) */
uint64_t pd_launch_count(const pd_batch* b);
/* name of the tick kernel this batch dispatches to ("k_tick_quad": <= 20480 envs, "k_tick": larger batches) */
const char* pd_tick_kernel(const pd_batch* b);
/* the exact kernel instance: "k_tick_quad<2>" / "<4>" / "<8>" (cars per warp; "<4,strut,dwb>" ... for the double-wishbone cars), "k_tick" (128
 * registers: batches beyond 37888 envs), "k_tick/255" (255 registers: thread-per-car batches that fit one wave at 4 blocks per SM), "k_tick<strut,dwb>" /
 * "k_tick<dwb,dwb>"; the parity tests run on every one */
const char* pd_tick_kernel_instance(const pd_batch* b);
/* suspension topology of the loaded car: (front == DWB) * 2 + (rear == DWB)  (Car.cpp:74-117: SuspensionStrut / SuspensionDW /
 * SuspensionAxle chosen from suspensions.ini [FRONT] / [REAR] TYPE) */
int pd_topology(const pd_batch* b);
/* the ray caster's triangle tree (general rays: Track::computeFatPoints' traces, pd_raycast; the reference's rays go through ODE / OPCODE,
 * Physics/ODE/RayCasterODE.cpp): built on the host at load (median splits) or ON THE DEVICE as a linear BVH -- automatically for tracks of
 * 400 000 triangles and more, or with PD_DEVICE_BVH=1 in the environment.  *depth is reported for the device-built tree. */
int pd_bvh_info(const pd_batch* b, int* built_on_device, int* n_nodes, int* depth);

#ifdef __cplusplus
}
#endif
#endif /* PD_BATCH_H */
