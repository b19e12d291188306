/*
 * pyprojectd_module.cpp -- the Python module `PyProjectD` (pybind11), a literal drop-in for the simulator half of the reference's
 * module (src/PyProjectD/PyProjectD.cpp:515-641) on top of the C ABI of include/pd_batch.h (libpd_b200.so).
 *
 * Same module name, class names, attribute names and function signatures, so that pyprojectd/projectd_env.py runs on it
 * UNMODIFIED (tests/test_pyprojectd_dropin.py does exactly that).  Error behaviour of the reference: nothing throws into
 * Python, failures are logged and give -1 (ids) or a silent no-op on bad ids (PyProjectD.cpp:100-136).
 *
 * One simulator = one CUDA batch.  createSimulator(basePath, numEnvs = 1, device = 0) has two extra optional arguments; with the
 * defaults it is the reference's call (one car per simulator, projectd_env.py:118-121).  `carId` is the env index inside the batch.
 * Batched extensions (not in the reference): setCarControlsBatch, setActionsBatch, envStep, getObsDLPack, getObsPtr.
 * Viewer / playground functions are stubs that report "not initialised" (the renderer is outside the hot path).
 */
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <pybind11/numpy.h>
#include <array>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/pd_batch.h"

namespace py = pybind11;

/* value types of the reference's bindings (Core/Math.h:91-207, Car/CarControls.h:9-20, Car/CarState.h:11-56; pack 4) */
struct vec3f { float x = 0, y = 0, z = 0; };
struct mat44f { float M11 = 0, M12 = 0, M13 = 0, M14 = 0, M21 = 0, M22 = 0, M23 = 0, M24 = 0, M31 = 0, M32 = 0, M33 = 0, M34 = 0, M41 = 0, M42 = 0, M43 = 0, M44 = 0; };
#pragma pack(push, 4)
struct CarControls {
    float steer = 0, clutch = 0, brake = 0, handBrake = 0, gas = 0;
    int8_t isShifterSupported = 1, requestedGearIndex = -1, gearUp = 0, gearDn = 0;
};
struct CarState {
    int32_t carId = 0, simId = 0; float timestamp = 0;
    CarControls controls;
    int32_t collisionFlag = 0, outOfTrackFlag = 0, trackPointId = 0; float lastTrackPointTimestamp = 0, trackLocation = 0, bodyVsTrack = 0, velocityVsTrack = 0;
    float engineRPM = 0, speedMS = 0; int32_t gear = 0, gearGrinding = 0;
    mat44f bodyMatrix; vec3f bodyPos, bodyEuler, accG, velocity, localVelocity, angularVelocity, localAngularVelocity;
    std::array<mat44f, 4> hubMatrix; std::array<vec3f, 4> tyreContacts;
    std::array<float, 4> tyreLoad, tyreAngularSpeed, tyreSlipRatio, tyreNdSlip;
    std::array<float, 10> probes; std::array<float, 5> lookAhead;
    float stepReward = 0, totalReward = 0;
};
#pragma pack(pop)
static_assert(sizeof(CarControls) == 24 && sizeof(CarState) == 664, "layouts of Car/CarControls.h and Car/CarState.h");

struct Sim {
    std::string base, track;
    int nEnvs = 1, device = 0;
    pd_batch* h = nullptr;
    bool tpCollision = false, tpBadLoc = false; int tpMode = 0;
    std::vector<float> ctl; std::vector<int8_t> gears;      /* host mirror of the controls: setCarControls edits one row */
    ~Sim() { if (h) pd_destroy(h); }
};
static std::map<int, std::shared_ptr<Sim>> g_sims;
static std::mutex g_mux;
static int g_nextId = 0;
static uint64_t g_seed = 0;
static std::string g_logFile;

static void logf(const char* fmt, ...) {
    char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    FILE* f = g_logFile.empty() ? nullptr : fopen(g_logFile.c_str(), "a");
    if (f) { fprintf(f, "[PyProjectD/b200] %s\n", buf); fclose(f); } else fprintf(stderr, "[PyProjectD/b200] %s\n", buf);
}
static std::shared_ptr<Sim> getSim(int simId) { std::lock_guard<std::mutex> l(g_mux); auto it = g_sims.find(simId); return it == g_sims.end() ? nullptr : it->second; }
static std::shared_ptr<Sim> getCar(int simId, int carId) { auto s = getSim(simId); return (s && s->h && carId >= 0 && carId < s->nEnvs) ? s : nullptr; }
static bool ok(const std::shared_ptr<Sim>& s, int rc, const char* what) { if (rc != PD_OK) { logf("%s: %s", what, pd_last_error(s->h)); return false; } return true; }
static std::vector<uint8_t> oneHot(int n, int i) { std::vector<uint8_t> m((size_t)n, 0); m[(size_t)i] = 1; return m; }

/* ---- PyProjectD.cpp:50-68 ---- */
static void setSeed(unsigned int seed) {
    g_seed = seed;
    std::lock_guard<std::mutex> l(g_mux);
    for (auto& kv : g_sims) if (kv.second->h) pd_set_seed(kv.second->h, g_seed, 0);
}
static void setLogFile(const std::string& path, bool overwrite) { g_logFile = path; if (overwrite) { FILE* f = fopen(path.c_str(), "w"); if (f) fclose(f); } }
static void clearLogFile() { if (!g_logFile.empty()) { FILE* f = fopen(g_logFile.c_str(), "w"); if (f) fclose(f); } }
static void writeLog(const std::string& msg) { logf("%s", msg.c_str()); }

/* ---- simulators (PyProjectD.cpp:111-203) ---- */
static int createSimulator(const std::string& basePath, int numEnvs, int device) {
    if (numEnvs <= 0) { logf("createSimulator: numEnvs must be positive"); return -1; }
    auto s = std::make_shared<Sim>(); s->base = basePath; s->nEnvs = numEnvs; s->device = device;
    std::lock_guard<std::mutex> l(g_mux);
    const int id = g_nextId++; g_sims[id] = s; return id;
}
static void destroySimulator(int simId) { std::shared_ptr<Sim> s; { std::lock_guard<std::mutex> l(g_mux); auto it = g_sims.find(simId); if (it != g_sims.end()) { s = it->second; g_sims.erase(it); } } }
static void loadTrack(int simId, const std::string& name) { if (auto s = getSim(simId)) s->track = name; }
static void unloadTrack(int simId) { if (auto s = getSim(simId)) { if (s->h) { pd_destroy(s->h); s->h = nullptr; } s->track.clear(); } }
static int addCar(int simId, const std::string& model) {
    auto s = getSim(simId);
    if (!s || s->track.empty()) { logf("addCar: no such simulator or no track loaded"); return -1; }
    if (s->h) { logf("addCar: this build holds one car model per simulator (its cars are the env batch)"); return -1; }
    if (pd_create(s->base.c_str(), s->track.c_str(), model.c_str(), s->nEnvs, s->device, &s->h) != PD_OK) { logf("addCar: %s", pd_last_error(nullptr)); s->h = nullptr; return -1; }
    pd_set_seed(s->h, g_seed, 0);
    s->ctl.assign((size_t)s->nEnvs * 5, 0.0f); s->gears.assign((size_t)s->nEnvs * 3, 0);
    for (int e = 0; e < s->nEnvs; ++e) s->gears[(size_t)e * 3] = -1;
    return 0;
}
static void removeCar(int simId, int) { if (auto s = getSim(simId)) { if (s->h) { pd_destroy(s->h); s->h = nullptr; } } }

/* ---- teleports (PyProjectD.cpp:245-295) ---- */
static void teleportCarToSpline(int simId, int carId, float distanceNorm) {
    if (auto s = getCar(simId, carId)) { std::vector<float> u((size_t)s->nEnvs, 0.0f); u[(size_t)carId] = distanceNorm; auto m = oneHot(s->nEnvs, carId); ok(s, pd_teleport_spline(s->h, m.data(), u.data()), "teleportCarToSpline"); }
}
static void teleportCarByMode(int simId, int carId, int mode) { if (auto s = getCar(simId, carId)) { auto m = oneHot(s->nEnvs, carId); ok(s, pd_teleport_mode(s->h, m.data(), mode), "teleportCarByMode"); } }
static void teleportCarToLocation(int simId, int carId, float, float, float) { if (getCar(simId, carId)) logf("teleportCarToLocation: not on the hot path of this build (use teleportCarToSpline / teleportCarByMode)"); }
static void teleportCarToPits(int simId, int carId, int) { if (auto s = getCar(simId, carId)) { auto m = oneHot(s->nEnvs, carId); ok(s, pd_teleport_mode(s->h, m.data(), PD_TELEPORT_START), "teleportCarToPits"); } }
static void setCarAutoTeleport(int simId, int carId, bool collision, bool badLoc, int mode) { if (auto s = getCar(simId, carId)) { s->tpCollision = collision; s->tpBadLoc = badLoc; s->tpMode = mode; } }

/* ---- controls / assists / tunes / scoring (PyProjectD.cpp:297-363) ---- */
static void setCarControls(int simId, int carId, bool smooth, const CarControls& c) {
    auto s = getCar(simId, carId); if (!s) return;
    float* r = &s->ctl[(size_t)carId * 5]; r[0] = c.steer; r[1] = c.clutch; r[2] = c.brake; r[3] = c.handBrake; r[4] = c.gas;
    int8_t* g = &s->gears[(size_t)carId * 3]; g[0] = c.requestedGearIndex; g[1] = c.gearUp; g[2] = c.gearDn;
    ok(s, pd_set_controls(s->h, s->ctl.data(), s->gears.data(), smooth ? 1 : 0, 0), "setCarControls");
}
static void setCarAssists(int simId, int carId, bool ac, bool as, bool ab) { if (auto s = getCar(simId, carId)) ok(s, pd_set_assists(s->h, ac, as, ab), "setCarAssists"); }
static void setCarTune(int simId, int carId, const std::string& name, float v) { if (auto s = getCar(simId, carId)) ok(s, pd_set_tune(s->h, name.c_str(), v), "setCarTune"); }
static void setCarRawTune(int simId, int carId, const std::string& name, float v) { if (auto s = getCar(simId, carId)) ok(s, pd_set_raw_tune(s->h, name.c_str(), v), "setCarRawTune"); }
static void setScoringVar(int simId, int carId, const std::string& name, float v) { if (auto s = getCar(simId, carId)) ok(s, pd_set_scoring_var(s->h, name.c_str(), v), "setScoringVar"); }
static float getScoringVar(int simId, int carId, const std::string& name) { auto s = getCar(simId, carId); return s ? pd_get_scoring_var(s->h, name.c_str()) : 0.0f; }

/* ---- stepping (PyProjectD.cpp:160-180): Simulator::step for every env; the auto-teleport of Car::postStep (Car.cpp:700-712) ---- */
static void stepSimulator(int simId, double dt) {
    auto s = getSim(simId); if (!s || !s->h) return;
    if (!ok(s, pd_step(s->h, (float)dt, 1), "stepSimulator")) return;
    if (s->tpCollision || s->tpBadLoc) {
        std::vector<int32_t> flags((size_t)s->nEnvs);
        if (!ok(s, pd_get_rewards(s->h, nullptr, nullptr, flags.data()), "stepSimulator")) return;
        std::vector<uint8_t> m((size_t)s->nEnvs, 0); bool any = false;
        for (int e = 0; e < s->nEnvs; ++e) if ((s->tpCollision && (flags[(size_t)e] & 1)) || (s->tpBadLoc && (flags[(size_t)e] & 2))) { m[(size_t)e] = 1; any = true; }
        if (any) pd_teleport_mode(s->h, m.data(), s->tpMode);
    }
}
static void getCarState(int simId, int carId, CarState& st) {
    auto s = getCar(simId, carId); if (!s) return;
    if (ok(s, pd_get_car_state(s->h, carId, &st), "getCarState")) { st.simId = simId; st.carId = carId; }
}

/* ---- batched extensions ---- */
static void setCarControlsBatch(int simId, py::array_t<float, py::array::c_style | py::array::forcecast> ctl, py::object gears, bool smooth) {
    auto s = getCar(simId, 0); if (!s) return;
    if (ctl.size() != (py::ssize_t)s->nEnvs * 5) { logf("setCarControlsBatch: controls must be [numEnvs, 5]"); return; }
    std::copy(ctl.data(), ctl.data() + ctl.size(), s->ctl.begin());
    if (!gears.is_none()) {
        auto g = gears.cast<py::array_t<int8_t, py::array::c_style | py::array::forcecast>>();
        if (g.size() != (py::ssize_t)s->nEnvs * 3) { logf("setCarControlsBatch: gears must be [numEnvs, 3]"); return; }
        std::copy(g.data(), g.data() + g.size(), s->gears.begin());
    }
    ok(s, pd_set_controls(s->h, s->ctl.data(), s->gears.data(), smooth ? 1 : 0, 0), "setCarControlsBatch");
}
static void setActionsBatch(int simId, py::array_t<float, py::array::c_style | py::array::forcecast> act) {
    auto s = getCar(simId, 0); if (!s) return;
    if (act.size() != (py::ssize_t)s->nEnvs * 2) { logf("setActionsBatch: actions must be [numEnvs, 2]"); return; }
    ok(s, pd_set_actions(s->h, act.data(), 0), "setActionsBatch");
}
/* one vectorised ProjectDEnv.step on DEVICE buffers given as raw addresses (torch: tensor.data_ptr()) */
static int envStep(int simId, uintptr_t actionsDev, double dt, uintptr_t obsDev, uintptr_t rewardDev, uintptr_t doneDev) {
    auto s = getCar(simId, 0); if (!s) return -1;
    return ok(s, pd_env_step(s->h, (const float*)actionsDev, (float)dt, (float*)obsDev, (float*)rewardDev, (int32_t*)doneDev), "envStep") ? 0 : -1;
}
static py::object getObsDLPack(int simId) {
    auto s = getCar(simId, 0); if (!s) return py::none();
    pd_observe(s->h);
    void* t = pd_obs_dlpack(s->h);
    return t ? py::reinterpret_steal<py::object>(PyCapsule_New(t, "dltensor", nullptr)) : py::none();
}
static uintptr_t getObsPtr(int simId) { auto s = getCar(simId, 0); return s ? (uintptr_t)pd_obs_device_ptr(s->h) : 0; }
static uintptr_t getStream(int simId) { auto s = getCar(simId, 0); return s ? (uintptr_t)pd_stream(s->h) : 0; }
static uintptr_t getBatchHandle(int simId) { auto s = getCar(simId, 0); return s ? (uintptr_t)s->h : 0; }
static int getNumEnvs(int simId) { auto s = getSim(simId); return s ? s->nEnvs : 0; }

/* ---- viewer half of the reference module (PyProjectD.cpp:371-509): outside the hot path, kept as inert stubs so that env code
 *      which probes them (projectd_env.py:137-155) keeps working ---- */
static void initPlayground(const std::string&) { logf("initPlayground: the renderer is not part of this build"); }
static void noop() {}
static bool retFalse() { return false; }
static bool retTrue() { return true; }
static void shutAll() { std::map<int, std::shared_ptr<Sim>> tmp; { std::lock_guard<std::mutex> l(g_mux); tmp.swap(g_sims); } }

PYBIND11_MODULE(PyProjectD, m) {
    m.doc() = "PyProjectD (B200 batched Car::step core behind the reference's module interface)";
    py::class_<vec3f>(m, "vec3f").def(py::init<>()).def_readwrite("x", &vec3f::x).def_readwrite("y", &vec3f::y).def_readwrite("z", &vec3f::z);
    py::class_<mat44f>(m, "mat44f").def(py::init<>())
        .def_readwrite("M11", &mat44f::M11).def_readwrite("M12", &mat44f::M12).def_readwrite("M13", &mat44f::M13).def_readwrite("M14", &mat44f::M14)
        .def_readwrite("M21", &mat44f::M21).def_readwrite("M22", &mat44f::M22).def_readwrite("M23", &mat44f::M23).def_readwrite("M24", &mat44f::M24)
        .def_readwrite("M31", &mat44f::M31).def_readwrite("M32", &mat44f::M32).def_readwrite("M33", &mat44f::M33).def_readwrite("M34", &mat44f::M34)
        .def_readwrite("M41", &mat44f::M41).def_readwrite("M42", &mat44f::M42).def_readwrite("M43", &mat44f::M43).def_readwrite("M44", &mat44f::M44);
    py::class_<CarControls>(m, "CarControls").def(py::init<>())
        .def_readwrite("steer", &CarControls::steer).def_readwrite("clutch", &CarControls::clutch).def_readwrite("brake", &CarControls::brake)
        .def_readwrite("handBrake", &CarControls::handBrake).def_readwrite("gas", &CarControls::gas)
        .def_readwrite("isShifterSupported", &CarControls::isShifterSupported).def_readwrite("requestedGearIndex", &CarControls::requestedGearIndex)
        .def_readwrite("gearUp", &CarControls::gearUp).def_readwrite("gearDn", &CarControls::gearDn);
    py::class_<CarState>(m, "CarState").def(py::init<>())
        .def_readonly("carId", &CarState::carId).def_readonly("simId", &CarState::simId).def_readonly("timestamp", &CarState::timestamp)
        .def_readonly("controls", &CarState::controls)
        .def_readonly("collisionFlag", &CarState::collisionFlag).def_readonly("outOfTrackFlag", &CarState::outOfTrackFlag)
        .def_readonly("trackPointId", &CarState::trackPointId).def_readonly("lastTrackPointTimestamp", &CarState::lastTrackPointTimestamp)
        .def_readonly("trackLocation", &CarState::trackLocation).def_readonly("bodyVsTrack", &CarState::bodyVsTrack).def_readonly("velocityVsTrack", &CarState::velocityVsTrack)
        .def_readonly("engineRPM", &CarState::engineRPM).def_readonly("speedMS", &CarState::speedMS).def_readonly("gear", &CarState::gear).def_readonly("gearGrinding", &CarState::gearGrinding)
        .def_readonly("bodyMatrix", &CarState::bodyMatrix).def_readonly("bodyPos", &CarState::bodyPos).def_readonly("bodyEuler", &CarState::bodyEuler)
        .def_readonly("accG", &CarState::accG).def_readonly("velocity", &CarState::velocity).def_readonly("localVelocity", &CarState::localVelocity)
        .def_readonly("angularVelocity", &CarState::angularVelocity).def_readonly("localAngularVelocity", &CarState::localAngularVelocity)
        .def_readonly("hubMatrix", &CarState::hubMatrix).def_readonly("tyreContacts", &CarState::tyreContacts).def_readonly("tyreLoad", &CarState::tyreLoad)
        .def_readonly("tyreAngularSpeed", &CarState::tyreAngularSpeed).def_readonly("tyreSlipRatio", &CarState::tyreSlipRatio).def_readonly("tyreNdSlip", &CarState::tyreNdSlip)
        .def_readonly("probes", &CarState::probes).def_readonly("lookAhead", &CarState::lookAhead)
        .def_readonly("stepReward", &CarState::stepReward).def_readonly("totalReward", &CarState::totalReward);

    m.def("setSeed", &setSeed, "");
    m.def("setLogFile", &setLogFile, "", py::arg("path"), py::arg("overwrite") = true);
    m.def("clearLogFile", &clearLogFile, "");
    m.def("writeLog", &writeLog, "");
    m.def("createSimulator", &createSimulator, "", py::arg("basePath"), py::arg("numEnvs") = 1, py::arg("device") = 0);
    m.def("destroySimulator", &destroySimulator, "");
    m.def("stepSimulator", &stepSimulator, "", py::arg("simId"), py::arg("dt") = 1.0 / 333.0);
    m.def("loadTrack", &loadTrack, "");
    m.def("unloadTrack", &unloadTrack, "");
    m.def("addCar", &addCar, "");
    m.def("removeCar", &removeCar, "");
    m.def("teleportCarToLocation", &teleportCarToLocation, "");
    m.def("teleportCarToPits", &teleportCarToPits, "");
    m.def("teleportCarToSpline", &teleportCarToSpline, "");
    m.def("teleportCarByMode", &teleportCarByMode, "");
    m.def("setCarAutoTeleport", &setCarAutoTeleport, "", py::arg("simId"), py::arg("carId"), py::arg("collision"), py::arg("badLoc"), py::arg("teleportMode") = 0);
    m.def("setCarControls", &setCarControls, "");
    m.def("setCarAssists", &setCarAssists, "");
    m.def("getCarState", &getCarState, "");
    m.def("setCarRawTune", &setCarRawTune, "");
    m.def("setCarTune", &setCarTune, "");
    m.def("setScoringVar", &setScoringVar, "");
    m.def("getScoringVar", &getScoringVar, "");
    /* batched extensions */
    m.def("setCarControlsBatch", &setCarControlsBatch, "", py::arg("simId"), py::arg("controls"), py::arg("gears") = py::none(), py::arg("smooth") = true);
    m.def("setActionsBatch", &setActionsBatch, "");
    m.def("envStep", &envStep, "", py::arg("simId"), py::arg("actionsDev"), py::arg("dt") = 1.0 / 333.0, py::arg("obsDev") = 0, py::arg("rewardDev") = 0, py::arg("doneDev") = 0);
    m.def("getObsDLPack", &getObsDLPack, "");
    m.def("getObsPtr", &getObsPtr, "");
    m.def("getStream", &getStream, "");
    m.def("getBatchHandle", &getBatchHandle, "");
    m.def("getNumEnvs", &getNumEnvs, "");
    /* viewer stubs */
    m.def("launchPlaygroundInOwnThread", [](const std::string&) { logf("launchPlaygroundInOwnThread: the renderer is not part of this build"); }, "");
    m.def("initPlayground", &initPlayground, "");
    m.def("shutPlayground", &noop, "");
    m.def("shutAll", &shutAll, "");
    m.def("tickPlayground", &noop, "");
    m.def("isPlaygroundInitialized", &retFalse, "");
    m.def("isPlaygroundExited", &retTrue, "");
    m.def("moveWindow", [](int, int) {}, "");
    m.def("resizeWindow", [](int, int) {}, "");
    m.def("setRenderHz", [](int, bool) {}, "");
    m.def("setActiveSimulator", [](int, bool) {}, "");
    m.def("setActiveCar", [](int, bool, bool) {}, "");
    m.def("getActiveSimulator", []() { return -1; }, "");
    m.def("getActiveCar", []() { return -1; }, "");
}
