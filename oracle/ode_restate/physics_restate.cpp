/*
 * oracle/ode_restate/physics_restate.cpp -- TEST INFRASTRUCTURE (oracle only).
 *
 * An IPhysicsEngine back-end (Physics/IPhysicsEngine.h:10-28) for the reference's UNMODIFIED
 * Car/Sim/Core sources, standing in for Physics/ODE/*.cpp whose third-party dependency (ODE 0.16.3)
 * is absent from /root/reference.  It replaces Physics/PhysicsFactory.cpp:6-9 and mirrors, call for
 * call, what the reference's wrappers ask ODE to do:
 *   bodies      Physics/ODE/RigidBodyODE.cpp:5-270
 *   joints      Physics/ODE/JointODE.cpp:21-88   (ERP/CFM overrides: only DBall honours them)
 *   rays        Physics/ODE/PhysicsEngineODE.cpp:168-214, RayCasterODE.cpp:11-28
 *   world step  Physics/ODE/PhysicsEngineODE.cpp:216-224  -> oder::world_step (ode_core.h)
 * The arithmetic is in ode_core.h (restated ODE).  Ray-vs-trimesh follows OPCODE's culling
 * Moller-Trumbore test (OPC_RayTriOverlap.h) and ODE's dCollideRTL contact construction
 * (collision_trimesh_ray.cpp): t < length, det > 1e-6 with back-face culling, contact normal =
 * normalised (v1-v0)x(v2-v0) after ODE's g1/g2 swap, minimum depth across meshes.  DEVIATION
 * (documented): within one mesh the closest stabbed triangle is returned, where ODE
 * with FirstContact=1 returns the first triangle met by OPCODE's tree walk.
 * Collision DETECTION between the car colliders and the track (collisionStep / collisionNearCallback /
 * onCollision, PhysicsEngineODE.cpp:228-341) is restated in ode_collide.h: frame parity, category / mask
 * matching, the body-local normal filter of box-vs-trimesh contacts, and the collision callback.  The contact
 * JOINTS (the response) are not created, and the callback receives a zero normal so that
 * Car::onCollisionCallback (Car.cpp:921-1044) sets collisionFlag / lastCollisionTime but derives no damage
 * from a contact normal this restatement does not have for mesh-vs-mesh pairs.
 */
#include "Physics/PhysicsFactory.h"
#include "Physics/IPhysicsEngine.h"
#include "Core/Diag.h"
#include "ode_core.h"
#include "ode_collide.h"
#include <cfloat>

namespace D {

using oder::dReal;

struct RestateEngine;
typedef std::shared_ptr<RestateEngine> RestateEnginePtr;

struct TriMeshR : public ITriMesh {
    std::vector<TriMeshVertex> vertices;
    std::vector<TriMeshIndex> indices;
    void resize(size_t vc, size_t ic) override { vertices.resize(vc); indices.resize(ic); }
    TriMeshVertex* getVB() override { return vertices.data(); }
    size_t getVertexCount() override { return vertices.size(); }
    TriMeshIndex* getIB() override { return indices.data(); }
    size_t getIndexCount() override { return indices.size(); }
};

struct CollisionMeshR : public ICollisionObject {
    ITriMeshPtr trimesh;
    void* userPointer = nullptr;
    unsigned long category = 0, mask = 0;
    bool isDynamic = false;
    float bbMin[3], bbMax[3];
    /* spatial index only (plays the role of OPCODE's AABB tree; does not change results):
       uniform x-z grid of triangle lists */
    int gnx = 1, gnz = 1; float gcell = 1.0f;
    std::vector<std::vector<uint32_t>> cells;
    void buildGrid() {
        const float target = 1.5f;
        gnx = std::max(1, std::min(128, (int)ceilf((bbMax[0] - bbMin[0]) / target)));
        gnz = std::max(1, std::min(128, (int)ceilf((bbMax[2] - bbMin[2]) / target)));
        cells.assign((size_t)gnx * gnz, {});
        const TriMeshVertex* vb = trimesh->getVB(); const TriMeshIndex* ib = trimesh->getIB();
        const size_t nt = trimesh->getIndexCount() / 3;
        const float sx = gnx / std::max(1e-6f, bbMax[0] - bbMin[0]), sz = gnz / std::max(1e-6f, bbMax[2] - bbMin[2]);
        for (size_t t = 0; t < nt; ++t) {
            float x0 = FLT_MAX, x1 = -FLT_MAX, z0 = FLT_MAX, z1 = -FLT_MAX;
            for (int k = 0; k < 3; ++k) { const TriMeshVertex& v = vb[ib[t * 3 + k]]; x0 = std::min(x0, v.x); x1 = std::max(x1, v.x); z0 = std::min(z0, v.z); z1 = std::max(z1, v.z); }
            int ix0 = std::max(0, std::min(gnx - 1, (int)floorf((x0 - bbMin[0]) * sx) - 1)), ix1 = std::max(0, std::min(gnx - 1, (int)floorf((x1 - bbMin[0]) * sx) + 1));
            int iz0 = std::max(0, std::min(gnz - 1, (int)floorf((z0 - bbMin[2]) * sz) - 1)), iz1 = std::max(0, std::min(gnz - 1, (int)floorf((z1 - bbMin[2]) * sz) + 1));
            for (int iz = iz0; iz <= iz1; ++iz) for (int ix = ix0; ix <= ix1; ++ix) cells[(size_t)iz * gnx + ix].push_back((uint32_t)t);
        }
    }
    void setUserPointer(void* d) override { userPointer = d; }
    void* getUserPointer() override { return userPointer; }
    unsigned long getGroup() override { return category; }
    unsigned long getMask() override { return mask; }
};

static const bool kDebugColl = getenv("PDREF_DEBUG_COLL") != nullptr, kDebugPairs = getenv("PDREF_DEBUG_PAIRS") != nullptr;

struct BoxColliderR { vec3f centre, size; unsigned long category, mask; };
struct MeshColliderR { std::shared_ptr<CollisionMeshR> mesh; float offR[9]; float offP[3]; };   /* geom offset: body-local = offR * v + offP */

struct RigidBodyR : public IRigidBody {
    oder::Body b;
    RestateEngine* core;
    std::vector<BoxColliderR> boxColliders;
    std::vector<MeshColliderR> meshColliders;
    explicit RigidBodyR(RestateEngine* c) : core(c) {}
    void setEnabled(bool) override {}
    bool isEnabled() override { return true; }
    void setAutoDisable(bool) override {}
    void stop() override { for (int i = 0; i < 3; ++i) { b.lvel[i] = 0; b.avel[i] = 0; b.facc[i] = 0; b.tacc[i] = 0; } }

    void setMassBox(float m, float x, float y, float z) override { oder::body_set_mass_box(b, m, x, y, z); }
    float getMass() override { return b.mass; }
    void setMassExplicitInertia(float, float, float, float) override { SHOULD_NOT_REACH_FATAL; /* not used by the demo car */ }
    vec3f getLocalInertia() override { return vec3f(b.I[0], b.I[1], b.I[2]); }

    vec3f localToWorld(const vec3f& p) override { dReal r[3]; oder::body_rel_point_pos(b, &p.x, r); return vec3f(r); }
    vec3f worldToLocal(const vec3f& p) override { dReal r[3]; oder::body_pos_rel_point(b, &p.x, r); return vec3f(r); }
    vec3f localToWorldNormal(const vec3f& p) override { dReal r[3]; oder::mul0_331(r, b.R, &p.x); return vec3f(r); }
    vec3f worldToLocalNormal(const vec3f& p) override { dReal r[3]; oder::mul1_331(r, b.R, &p.x); return vec3f(r); }

    void setPosition(const vec3f& pos) override { b.pos[0] = pos.x; b.pos[1] = pos.y; b.pos[2] = pos.z; }
    vec3f getPosition(float) override { return vec3f(b.pos); }
    void setRotation(const mat44f& m) override {
        /* RigidBodyODE.cpp:143-156: r[0]=M11 r[1]=M21 r[2]=M31 / r[4]=M12 ... (ODE rows = mat44f columns) */
        dReal R[9] = {m.M11, m.M21, m.M31, m.M12, m.M22, m.M32, m.M13, m.M23, m.M33};
        oder::body_set_rotation(b, R);
    }
    mat44f getWorldMatrix(float) override {
        mat44f m; const dReal* r = b.R;
        m.M11 = r[0]; m.M12 = r[3]; m.M13 = r[6]; m.M14 = 0;
        m.M21 = r[1]; m.M22 = r[4]; m.M23 = r[7]; m.M24 = 0;
        m.M31 = r[2]; m.M32 = r[5]; m.M33 = r[8]; m.M34 = 0;
        m.M41 = b.pos[0]; m.M42 = b.pos[1]; m.M43 = b.pos[2]; m.M44 = 1.0f;
        return m;
    }

    void setVelocity(const vec3f& v) override { b.lvel[0] = v.x; b.lvel[1] = v.y; b.lvel[2] = v.z; }
    vec3f getVelocity() override { dReal z[3] = {0, 0, 0}, r[3]; oder::body_rel_point_vel(b, z, r); return vec3f(r); }
    vec3f getLocalVelocity() override { return worldToLocalNormal(getVelocity()); }
    vec3f getPointVelocity(const vec3f& p) override { dReal r[3]; oder::body_point_vel(b, &p.x, r); return vec3f(r); }
    vec3f getLocalPointVelocity(const vec3f& p) override { dReal r[3]; oder::body_rel_point_vel(b, &p.x, r); return vec3f(r); }

    void setAngularVelocity(const vec3f& v) override { b.avel[0] = v.x; b.avel[1] = v.y; b.avel[2] = v.z; }
    vec3f getAngularVelocity() override { return vec3f(b.avel); }
    vec3f getLocalAngularVelocity() override { return worldToLocalNormal(getAngularVelocity()); }

    void addForceAtPos(const vec3f& f, const vec3f& p) override { oder::body_add_force_at_pos(b, &f.x, &p.x); }
    void addForceAtLocalPos(const vec3f& f, const vec3f& p) override { oder::body_add_force_at_rel_pos(b, &f.x, &p.x); }
    void addLocalForce(const vec3f& f) override { dReal z[3] = {0, 0, 0}, fw[3]; oder::mul0_331(fw, b.R, &f.x); oder::body_add_force_at_rel_pos(b, fw, z); }
    void addLocalForceAtPos(const vec3f& f, const vec3f& p) override { dReal fw[3]; oder::mul0_331(fw, b.R, &f.x); oder::body_add_force_at_pos(b, fw, &p.x); }
    void addLocalForceAtLocalPos(const vec3f& f, const vec3f& p) override { dReal fw[3]; oder::mul0_331(fw, b.R, &f.x); oder::body_add_force_at_rel_pos(b, fw, &p.x); }
    void addTorque(const vec3f& t) override { b.tacc[0] += t.x; b.tacc[1] += t.y; b.tacc[2] += t.z; }
    void addLocalTorque(const vec3f& t) override { dReal tw[3]; oder::mul0_331(tw, b.R, &t.x); b.tacc[0] += tw[0]; b.tacc[1] += tw[1]; b.tacc[2] += tw[2]; }

    /* RigidBodyODE.cpp:274-292: dCreateBox(size) attached to the body, offset position = pos, rotation = the body's */
    void addBoxCollider(const vec3f& pos, const vec3f& size, unsigned int, unsigned int category, unsigned long mask) override {
        boxColliders.push_back(BoxColliderR{pos, size, (unsigned long)category, mask});
    }
    void addMeshCollider(ITriMeshPtr trimesh, const mat44f& offset, unsigned int spaceId, unsigned long category, unsigned long mask) override;
};

struct JointR : public IJoint {
    oder::Joint j;
    float distance = 0; /* DistanceJointODE::distance */
    void setERPCFM(float erp, float cfm) override {
        /* JointODE.cpp:59-75: only Slider and DBall override.  Slider forwards to dJointSetSliderParam
           (dParamERP is not a limit-motor parameter; dParamCFM only feeds the absent limit/motor row)
           -> no effect on the 5 slider rows.  DBall: dJointSetDBallParam honours both. */
        if (j.type == oder::J_DBALL) { if (erp > 0.0f) j.erp = erp; if (cfm > 0.0f) j.cfm = cfm; }
    }
    void reseatDistanceJointLocal(const vec3f& p1, const vec3f& p2) override {
        if (j.type != oder::J_DBALL) return;
        /* JointODE.cpp:77-88: local -> world -> SetDBallAnchor (world -> local) -> SetDBallDistance */
        dReal w1[3], w2[3];
        oder::body_rel_point_pos(*j.b0, &p1.x, w1); oder::body_rel_point_pos(*j.b1, &p2.x, w2);
        oder::joint_set_dball_anchor1(j, w1); oder::joint_set_dball_anchor2(j, w2);
        j.targetDistance = distance;
    }
};

struct RayCasterR : public IRayCaster {
    RestateEngine* core; float length;
    RayCasterR(RestateEngine* c, float l) : core(c), length(l) {}
    RayCastHit rayCast(const vec3f& pos, const vec3f& dir) override;
};

struct RestateEngine : public IPhysicsEngine, std::enable_shared_from_this<RestateEngine> {
    oder::World world;
    std::vector<std::shared_ptr<RigidBodyR>> bodies;
    std::vector<std::shared_ptr<JointR>> joints;
    std::vector<std::shared_ptr<CollisionMeshR>> staticMeshes;
    ICollisionCallback* cb = nullptr;
    unsigned long long rayCount = 0, rayTriTests = 0;

    RestateEngine() {
        /* PhysicsEngineODE.cpp:23-28 */
        world.gravity[0] = 0; world.gravity[1] = -9.80665f; world.gravity[2] = 0;
        world.erp = 0.3f; world.cfm = 1.0e-7f;
    }
    oder::Body* B(IRigidBodyPtr rb) { return &std::dynamic_pointer_cast<RigidBodyR>(rb)->b; }

    IRigidBodyPtr createRigidBody() override {
        auto p = std::make_shared<RigidBodyR>(this);
        bodies.push_back(p); world.bodies.push_back(&p->b);
        return p;
    }
    std::shared_ptr<JointR> mk(oder::JointType t, IRigidBodyPtr a, IRigidBodyPtr b) {
        auto p = std::make_shared<JointR>();
        p->j.type = t; p->j.b0 = B(a); p->j.b1 = B(b);
        p->j.erp = world.erp; p->j.cfm = world.cfm;  /* joints copy the world defaults at creation */
        joints.push_back(p); world.joints.push_back(&p->j);
        return p;
    }
    IJointPtr createFixedJoint(IRigidBodyPtr a, IRigidBodyPtr b) override { auto p = mk(oder::J_FIXED, a, b); oder::joint_set_fixed(p->j); return p; }
    IJointPtr createBallJoint(IRigidBodyPtr a, IRigidBodyPtr b, const vec3f& pos) override { auto p = mk(oder::J_BALL, a, b); oder::joint_set_ball_anchor(p->j, &pos.x); return p; }
    IJointPtr createSliderJoint(IRigidBodyPtr a, IRigidBodyPtr b, const vec3f& axis) override { auto p = mk(oder::J_SLIDER, a, b); oder::joint_set_slider_axis(p->j, &axis.x); return p; }
    IJointPtr createDistanceJoint(IRigidBodyPtr a, IRigidBodyPtr b, const vec3f& p1, const vec3f& p2) override {
        auto p = mk(oder::J_DBALL, a, b);
        oder::joint_set_dball_anchor1(p->j, &p1.x); oder::joint_set_dball_anchor2(p->j, &p2.x);
        p->distance = p->j.targetDistance;
        return p;
    }
    IJointPtr createBumpJoint(IRigidBodyPtr, IRigidBodyPtr, const vec3f&, float, float) override { return nullptr; }

    ITriMeshPtr createTriMesh() override { return std::make_shared<TriMeshR>(); }
    ICollisionObjectPtr createCollider(ITriMeshPtr tm, bool isDynamic, unsigned int, unsigned long category, unsigned long mask) override {
        auto p = std::make_shared<CollisionMeshR>();
        p->trimesh = tm; p->category = category; p->mask = mask; p->isDynamic = isDynamic;
        for (int k = 0; k < 3; ++k) { p->bbMin[k] = FLT_MAX; p->bbMax[k] = -FLT_MAX; }
        auto* vb = tm->getVB();
        for (size_t i = 0; i < tm->getVertexCount(); ++i) {
            const float* v = &vb[i].x;
            for (int k = 0; k < 3; ++k) { p->bbMin[k] = std::min(p->bbMin[k], v[k]); p->bbMax[k] = std::max(p->bbMax[k], v[k]); }
        }
        if (!isDynamic) { p->buildGrid(); staticMeshes.push_back(p); }
        return p;
    }
    IRayCasterPtr createRayCaster(float length) override { return std::make_shared<RayCasterR>(this, length); }
    void setCollisionCallback(ICollisionCallback* c) override { cb = c; }

    RayCastHit rayCastImpl(const vec3f& o, const vec3f& d, float length) {
        RayCastHit hit;
        float best = -1.0f; vec3f bestN; CollisionMeshR* bestMesh = nullptr;
        float e[3] = {o.x + d.x * length, o.y + d.y * length, o.z + d.z * length};
        float rmin[3] = {std::min(o.x, e[0]), std::min(o.y, e[1]), std::min(o.z, e[2])};
        float rmax[3] = {std::max(o.x, e[0]), std::max(o.y, e[1]), std::max(o.z, e[2])};
        ++rayCount;
        for (auto& mp : staticMeshes) {
            CollisionMeshR& m = *mp;
            if (rmin[0] > m.bbMax[0] || rmax[0] < m.bbMin[0] || rmin[1] > m.bbMax[1] || rmax[1] < m.bbMin[1] || rmin[2] > m.bbMax[2] || rmax[2] < m.bbMin[2]) continue;
            const TriMeshVertex* vb = m.trimesh->getVB(); const TriMeshIndex* ib = m.trimesh->getIB();
            float meshBest = -1.0f; vec3f meshN;
            const float sx = m.gnx / std::max(1e-6f, m.bbMax[0] - m.bbMin[0]), sz = m.gnz / std::max(1e-6f, m.bbMax[2] - m.bbMin[2]);
            int ix0 = std::max(0, std::min(m.gnx - 1, (int)floorf((rmin[0] - m.bbMin[0]) * sx))), ix1 = std::max(0, std::min(m.gnx - 1, (int)floorf((rmax[0] - m.bbMin[0]) * sx)));
            int iz0 = std::max(0, std::min(m.gnz - 1, (int)floorf((rmin[2] - m.bbMin[2]) * sz))), iz1 = std::max(0, std::min(m.gnz - 1, (int)floorf((rmax[2] - m.bbMin[2]) * sz)));
            for (int iz = iz0; iz <= iz1; ++iz) for (int ix = ix0; ix <= ix1; ++ix)
            for (uint32_t t : m.cells[(size_t)iz * m.gnx + ix]) {
                const TriMeshVertex& v0 = vb[ib[t * 3]]; const TriMeshVertex& v1 = vb[ib[t * 3 + 1]]; const TriMeshVertex& v2 = vb[ib[t * 3 + 2]];
                ++rayTriTests;
                vec3f e1(v1.x - v0.x, v1.y - v0.y, v1.z - v0.z), e2(v2.x - v0.x, v2.y - v0.y, v2.z - v0.z);
                vec3f pvec = d.cross(e2);
                float det = e1 * pvec;
                if (det < 0.000001f) continue;                 /* LOCAL_EPSILON, culling on */
                vec3f tvec(o.x - v0.x, o.y - v0.y, o.z - v0.z);
                float u = tvec * pvec;
                if (u < 0.0f || u > det) continue;
                vec3f qvec = tvec.cross(e1);
                float v = d * qvec;
                if (v < 0.0f || u + v > det) continue;
                float dist = e2 * qvec;
                if (dist < 0.0f) continue;
                dist *= (1.0f / det);
                if (!(dist < length)) continue;
                if (meshBest < 0.0f || dist < meshBest) { meshBest = dist; meshN = e1.cross(e2); }
            }
            if (meshBest >= 0.0f && (best < 0.0f || best > meshBest)) { best = meshBest; bestN = meshN; bestMesh = &m; }
        }
        if (best >= 0.0f) {
            hit.pos = vec3f(o.x + d.x * best, o.y + d.y * best, o.z + d.z * best);
            hit.normal = bestN.get_norm();
            hit.collisionObject = bestMesh;
            hit.hasContact = true;
        }
        return hit;
    }
    RayCastHit rayCast(const vec3f& pos, const vec3f& dir, float length) override { return rayCastImpl(pos, dir, length); }
    RayCastHit rayCast(const vec3f& pos, const vec3f& dir, IRayCasterPtr ray) override {
        return rayCastImpl(pos, dir, std::dynamic_pointer_cast<RayCasterR>(ray)->length);
    }

    /* ---- collisionStep (PhysicsEngineODE.cpp:228-244): odd frames dynamic x static, even frames dynamic x dynamic ---- */
    unsigned int currentFrame = 0;
    unsigned long long collisionPairs = 0, localBoundHits = 0;
    static oder::CV3 toWorld(const oder::Body& b, float x, float y, float z) {
        return oder::cv((b.R[0] * x + b.R[1] * y + b.R[2] * z) + b.pos[0], (b.R[3] * x + b.R[4] * y + b.R[5] * z) + b.pos[1], (b.R[6] * x + b.R[7] * y + b.R[8] * z) + b.pos[2]);
    }
    void fire(RigidBodyR* rb, ICollisionObject* shape0, CollisionMeshR* other, oder::CV3 pos) {
        /* onCollision (PhysicsEngineODE.cpp:284-341) -> collisionCallback; normal 0: see the header */
        if (cb) cb->onCollisionCallback(rb, shape0, nullptr, other, vec3f(0, 0, 0), vec3f(pos.x, pos.y, pos.z), 0.0f);
    }
    /* candidate triangles of a static mesh near the world box [lo, hi]: the mesh's x-z grid (a spatial index only, it plays
       the role of OPCODE's tree), every triangle once */
    std::vector<uint32_t> cand;
    std::vector<oder::CV3> hl;
    void gather(CollisionMeshR& m, const float* lo, const float* hi) {
        cand.clear();
        const float sx = m.gnx / std::max(1e-6f, m.bbMax[0] - m.bbMin[0]), sz = m.gnz / std::max(1e-6f, m.bbMax[2] - m.bbMin[2]);
        const int ix0 = std::max(0, std::min(m.gnx - 1, (int)floorf((lo[0] - m.bbMin[0]) * sx))), ix1 = std::max(0, std::min(m.gnx - 1, (int)floorf((hi[0] - m.bbMin[0]) * sx)));
        const int iz0 = std::max(0, std::min(m.gnz - 1, (int)floorf((lo[2] - m.bbMin[2]) * sz))), iz1 = std::max(0, std::min(m.gnz - 1, (int)floorf((hi[2] - m.bbMin[2]) * sz)));
        for (int iz = iz0; iz <= iz1; ++iz) for (int ix = ix0; ix <= ix1; ++ix) for (uint32_t t : m.cells[(size_t)iz * m.gnx + ix]) cand.push_back(t);
        std::sort(cand.begin(), cand.end()); cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    }
    /* ---- contact generation (definition: header of ode_collide.h) + contact joints (PhysicsEngineODE.cpp:284-341) ---- */
    bool responseEnabled = true;
    std::vector<oder::ContactJoint> contactGroupDynamic;      /* created on odd frames, alive for that frame and the next one */
    void addContacts(RigidBodyR* rb, const oder::ContactPoint* cp, int n, ICollisionObject* shape0, CollisionMeshR* other) {
        for (int i = 0; i < n; ++i) {
            oder::ContactJoint cj; cj.b0 = &rb->b;
            cj.pos[0] = cp[i].pos.x; cj.pos[1] = cp[i].pos.y; cj.pos[2] = cp[i].pos.z;
            cj.normal[0] = cp[i].normal.x; cj.normal[1] = cp[i].normal.y; cj.normal[2] = cp[i].normal.z; cj.depth = cp[i].depth;
            if (cp[i].kind == 0) { cj.mu = 0.1f; cj.bounce = 0.0f; cj.soft_cfm = 0.000952380942f; cj.soft_erp = 0.714285731f; }   /* box <-> trimesh: mode 28700 */
            else { cj.mu = 0.25f; cj.bounce = 0.01f; cj.soft_cfm = 0.0001f; cj.soft_erp = -1.0f; }                                /* anything else: mode 28692 */
            contactGroupDynamic.push_back(cj);
            if (cb) cb->onCollisionCallback(rb, shape0, nullptr, other, vec3f(cj.normal[0], cj.normal[1], cj.normal[2]), vec3f(cj.pos[0], cj.pos[1], cj.pos[2]), cj.depth);
        }
    }
    void collideBodyStatic(RigidBodyR* rb) {
        const oder::Body& b = rb->b;
        const oder::CV3 A[3] = {oder::cv(b.R[0], b.R[3], b.R[6]), oder::cv(b.R[1], b.R[4], b.R[7]), oder::cv(b.R[2], b.R[5], b.R[8])};
        for (const BoxColliderR& bx : rb->boxColliders) {
            const oder::CV3 c = toWorld(b, bx.centre.x, bx.centre.y, bx.centre.z);
            const float h[3] = {bx.size.x * 0.5f, bx.size.y * 0.5f, bx.size.z * 0.5f};
            float ext[3];
            for (int k = 0; k < 3; ++k) ext[k] = h[0] * fabsf(b.R[k * 3 + 0]) + h[1] * fabsf(b.R[k * 3 + 1]) + h[2] * fabsf(b.R[k * 3 + 2]);
            const float lo[3] = {c.x - ext[0], c.y - ext[1], c.z - ext[2]}, hi[3] = {c.x + ext[0], c.y + ext[1], c.z + ext[2]};
            oder::CV3 corner[8];
            for (int k = 0; k < 8; ++k) {
                const float sx = (k & 1) ? 1.0f : -1.0f, sy = (k & 2) ? 1.0f : -1.0f, sz = (k & 4) ? 1.0f : -1.0f;
                corner[k] = toWorld(b, bx.centre.x + sx * h[0], bx.centre.y + sy * h[1], bx.centre.z + sz * h[2]);
            }
            oder::ContactPoint slot[8]; bool used[8] = {false, false, false, false, false, false, false, false};
            bool anyAccepted = false; float fbDepth = -1.0f; oder::CV3 fbN = oder::cv(0, 0, 0); CollisionMeshR* firstMesh = nullptr;
            for (auto& mp : staticMeshes) {
                CollisionMeshR& m = *mp;
                if (!((bx.category & m.mask) && (m.category & bx.mask))) continue;      /* collisionNearCallback bMatch */
                if (lo[0] > m.bbMax[0] || hi[0] < m.bbMin[0] || lo[1] > m.bbMax[1] || hi[1] < m.bbMin[1] || lo[2] > m.bbMax[2] || hi[2] < m.bbMin[2]) continue;
                const TriMeshVertex* vb = m.trimesh->getVB(); const TriMeshIndex* ib = m.trimesh->getIB();
                gather(m, lo, hi);
                for (uint32_t t : cand) {
                    const TriMeshVertex& a0 = vb[ib[t * 3]]; const TriMeshVertex& a1 = vb[ib[t * 3 + 1]]; const TriMeshVertex& a2 = vb[ib[t * 3 + 2]];
                    if (std::min(a0.x, std::min(a1.x, a2.x)) > hi[0] || std::max(a0.x, std::max(a1.x, a2.x)) < lo[0]) continue;
                    if (std::min(a0.y, std::min(a1.y, a2.y)) > hi[1] || std::max(a0.y, std::max(a1.y, a2.y)) < lo[1]) continue;
                    if (std::min(a0.z, std::min(a1.z, a2.z)) > hi[2] || std::max(a0.z, std::max(a1.z, a2.z)) < lo[2]) continue;
                    ++collisionPairs;
                    oder::CV3 n; float satDepth = 0.0f;
                    const oder::CV3 t0 = oder::cv(a0.x, a0.y, a0.z), t1 = oder::cv(a1.x, a1.y, a1.z), t2 = oder::cv(a2.x, a2.y, a2.z);
                    if (!oder::box_tri_contact(c, A, h, t0, t1, t2, n, &satDepth)) continue;
                    /* box vs trimesh: contacts whose body-local normal y < 0.9 are dropped (PhysicsEngineODE.cpp:309-318) */
                    const float locY = b.R[1] * n.x + b.R[4] * n.y + b.R[7] * n.z;
                    if (locY < 0.9f) continue;
                    if (kDebugColl) fprintf(stderr, "[oracle] box contact: tri %zu of mesh cat %lu, n=(%g %g %g) locY=%g, tri y=(%g %g %g) box c=(%g %g %g)\n", t, m.category, n.x, n.y, n.z, locY, a0.y, a1.y, a2.y, c.x, c.y, c.z);
                    fire(rb, nullptr, &m, c);
                    if (!responseEnabled) continue;
                    /* contact candidates of this accepted triangle: box corners below its plane, inside its prism */
                    if (!firstMesh) firstMesh = &m;
                    anyAccepted = true;
                    if (satDepth > fbDepth || (satDepth == fbDepth && (n.y > fbN.y || (n.y == fbN.y && (n.x > fbN.x || (n.x == fbN.x && n.z > fbN.z)))))) { fbDepth = satDepth; fbN = n; }
                    oder::CV3 N = oder::ccross(oder::csub(t1, t0), oder::csub(t2, t0));
                    const float len = sqrtf(oder::cdot(N, N));
                    if (!(len > 1e-12f)) continue;
                    const float inv = 1.0f / len;
                    oder::CV3 nt = oder::cv(N.x * inv, N.y * inv, N.z * inv);
                    if (oder::cdot(nt, n) < 0.0f) nt = oder::cv(-nt.x, -nt.y, -nt.z);          /* towards the box */
                    for (int k = 0; k < 8; ++k) {
                        const float sdist = oder::cdot(oder::csub(corner[k], t0), nt);
                        if (!(sdist < 0.0f)) continue;
                        if (!oder::projects_inside(corner[k], t0, t1, t2, N)) continue;
                        oder::contact_offer(slot, used, k, corner[k], nt, -sdist, 0);
                    }
                }
            }
            if (responseEnabled && anyAccepted) {
                oder::ContactPoint out[4];
                int nOut = oder::contact_select(slot, used, 8, out, 4);
                if (nOut == 0) {       /* no corner qualified: one contact at the support corner of the deepest accepted triangle */
                    int bestK = 0; float bestS = 3.4e38f;
                    for (int k = 0; k < 8; ++k) { const float sdot = oder::cdot(corner[k], fbN); if (sdot < bestS) { bestS = sdot; bestK = k; } }
                    out[0].pos = corner[bestK]; out[0].normal = fbN; out[0].depth = fbDepth; out[0].kind = 0; nOut = 1;
                }
                addContacts(rb, out, nOut, nullptr, firstMesh);
            }
        }
        for (const MeshColliderR& mc : rb->meshColliders) {
            /* mesh vs mesh in the MODEL space of the car's mesh (as OPCODE's tree-vs-tree query works): the hull keeps its
               body-local vertices (offset applied), every candidate wall triangle is brought into the chassis frame */
            CollisionMeshR& cm = *mc.mesh;
            const TriMeshVertex* cvb = cm.trimesh->getVB(); const TriMeshIndex* cib = cm.trimesh->getIB();
            const size_t nv = cm.trimesh->getVertexCount(), nct = cm.trimesh->getIndexCount() / 3;
            hl.resize(nv);
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            const float* r = mc.offR;
            for (size_t i = 0; i < nv; ++i) {
                hl[i] = oder::cv((r[0] * cvb[i].x + r[1] * cvb[i].y + r[2] * cvb[i].z) + mc.offP[0], (r[3] * cvb[i].x + r[4] * cvb[i].y + r[5] * cvb[i].z) + mc.offP[1], (r[6] * cvb[i].x + r[7] * cvb[i].y + r[8] * cvb[i].z) + mc.offP[2]);
                const oder::CV3 w = toWorld(b, hl[i].x, hl[i].y, hl[i].z);
                lo[0] = std::min(lo[0], w.x); hi[0] = std::max(hi[0], w.x); lo[1] = std::min(lo[1], w.y); hi[1] = std::max(hi[1], w.y); lo[2] = std::min(lo[2], w.z); hi[2] = std::max(hi[2], w.z);
            }
            auto toLocal = [&](const TriMeshVertex& v) {
                const float dx = v.x - b.pos[0], dy = v.y - b.pos[1], dz = v.z - b.pos[2];
                return oder::cv(b.R[0] * dx + b.R[3] * dy + b.R[6] * dz, b.R[1] * dx + b.R[4] * dy + b.R[7] * dz, b.R[2] * dx + b.R[5] * dy + b.R[8] * dz);
            };
            std::vector<oder::ContactPoint> vslot(nv); std::vector<char> vusedC(nv, 0);
            CollisionMeshR* firstWall = nullptr;
            for (auto& mp : staticMeshes) {
                CollisionMeshR& m = *mp;
                if (!((cm.category & m.mask) && (m.category & cm.mask))) continue;
                if (lo[0] > m.bbMax[0] || hi[0] < m.bbMin[0] || lo[1] > m.bbMax[1] || hi[1] < m.bbMin[1] || lo[2] > m.bbMax[2] || hi[2] < m.bbMin[2]) continue;
                const TriMeshVertex* vb = m.trimesh->getVB(); const TriMeshIndex* ib = m.trimesh->getIB();
                gather(m, lo, hi);
                bool hit = false;
                for (uint32_t t : cand) {
                    if (hit && !responseEnabled) break;
                    const TriMeshVertex& a0 = vb[ib[t * 3]]; const TriMeshVertex& a1 = vb[ib[t * 3 + 1]]; const TriMeshVertex& a2 = vb[ib[t * 3 + 2]];
                    if (std::min(a0.x, std::min(a1.x, a2.x)) > hi[0] || std::max(a0.x, std::max(a1.x, a2.x)) < lo[0]) continue;
                    if (std::min(a0.y, std::min(a1.y, a2.y)) > hi[1] || std::max(a0.y, std::max(a1.y, a2.y)) < lo[1]) continue;
                    if (std::min(a0.z, std::min(a1.z, a2.z)) > hi[2] || std::max(a0.z, std::max(a1.z, a2.z)) < lo[2]) continue;
                    const oder::CV3 b0 = toLocal(a0), b1 = toLocal(a1), b2 = toLocal(a2);
                    if (kDebugPairs) { float l0[3] = {std::min(b0.x, std::min(b1.x, b2.x)), std::min(b0.y, std::min(b1.y, b2.y)), std::min(b0.z, std::min(b1.z, b2.z))}, h0[3] = {std::max(b0.x, std::max(b1.x, b2.x)), std::max(b0.y, std::max(b1.y, b2.y)), std::max(b0.z, std::max(b1.z, b2.z))}; float hlo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hhi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX}; for (auto& q : hl) { hlo[0] = std::min(hlo[0], q.x); hhi[0] = std::max(hhi[0], q.x); hlo[1] = std::min(hlo[1], q.y); hhi[1] = std::max(hhi[1], q.y); hlo[2] = std::min(hlo[2], q.z); hhi[2] = std::max(hhi[2], q.z); } if (!(l0[0] > hhi[0] || h0[0] < hlo[0] || l0[1] > hhi[1] || h0[1] < hlo[1] || l0[2] > hhi[2] || h0[2] < hlo[2])) ++localBoundHits; }
                    bool triHit = false;
                    for (size_t k = 0; k < nct && !triHit; ++k) {
                        ++collisionPairs;
                        if (oder::tri_tri(hl[cib[k * 3]], hl[cib[k * 3 + 1]], hl[cib[k * 3 + 2]], b0, b1, b2)) triHit = true;
                    }
                    if (!triHit) continue;
                    hit = true;
                    if (!responseEnabled) continue;
                    if (!firstWall) firstWall = &m;
                    /* contact candidates of this wall triangle, in the chassis frame (the hull's model space): hull vertices behind its
                       plane (seen from the chassis origin) by less than 0.5 m, inside its prism */
                    oder::CV3 N = oder::ccross(oder::csub(b1, b0), oder::csub(b2, b0));
                    const float len = sqrtf(oder::cdot(N, N));
                    if (!(len > 1e-12f)) continue;
                    const float inv = 1.0f / len;
                    oder::CV3 nl = oder::cv(N.x * inv, N.y * inv, N.z * inv);
                    if (oder::cdot(oder::csub(oder::cv(0, 0, 0), b0), nl) < 0.0f) nl = oder::cv(-nl.x, -nl.y, -nl.z);      /* towards the chassis origin */
                    for (size_t j = 0; j < nv; ++j) {
                        const float sdist = oder::cdot(oder::csub(hl[j], b0), nl);
                        if (!(sdist < 0.0f) || !(sdist > -0.5f)) continue;
                        if (!oder::projects_inside(hl[j], b0, b1, b2, N)) continue;
                        const oder::CV3 pw = toWorld(b, hl[j].x, hl[j].y, hl[j].z);
                        const oder::CV3 nw = oder::cv(b.R[0] * nl.x + b.R[1] * nl.y + b.R[2] * nl.z, b.R[3] * nl.x + b.R[4] * nl.y + b.R[5] * nl.z, b.R[6] * nl.x + b.R[7] * nl.y + b.R[8] * nl.z);
                        bool u = vusedC[j] != 0;
                        oder::contact_offer(&vslot[j], &u, 0, pw, nw, -sdist, 1);
                        vusedC[j] = u ? 1 : 0;
                    }
                }
                if (hit && kDebugColl) fprintf(stderr, "[oracle] hull contact with mesh cat %lu\n", m.category);
                if (hit) fire(rb, &cm, &m, oder::cv(b.pos[0], b.pos[1], b.pos[2]));     /* one callback per mesh pair is enough for the flag */
            }
            if (responseEnabled && firstWall) {
                std::unique_ptr<bool[]> usedArr(new bool[nv ? nv : 1]); for (size_t j = 0; j < nv; ++j) usedArr[j] = vusedC[j] != 0;      /* every hull vertex is a slot (bundled cars: 50 .. 113 vertices) */
                oder::ContactPoint out[4];
                const int nOut = oder::contact_select(vslot.data(), usedArr.get(), (int)nv, out, 4);
                addContacts(rb, out, nOut, &cm, firstWall);
            }
        }
    }
    void collisionStep() {
        const unsigned long long pairs0 = collisionPairs;
        if (currentFrame & 1) { contactGroupDynamic.clear(); for (auto& rb : bodies) if (!rb->boxColliders.empty() || !rb->meshColliders.empty()) collideBodyStatic(rb.get()); }
        /* even frames: dynamic x dynamic -- one car per simulator here; its own box and mesh do not match each other's masks */
        if ((currentFrame & 1) && kDebugPairs) { fprintf(stderr, "[oracle] frame %u: %llu narrow-phase pairs %llu inbounds\n", currentFrame, collisionPairs - pairs0, localBoundHits); localBoundHits = 0; }
        currentFrame++;
    }

    void step(float dt) override {
        collisionStep();
        world.contacts = contactGroupDynamic;       /* both frames of the group's life (PhysicsEngineODE.cpp:228-244) */
        oder::world_step(world, dt);
    }
};

void RigidBodyR::addMeshCollider(ITriMeshPtr trimesh, const mat44f& offset, unsigned int spaceId, unsigned long category, unsigned long mask) {
    /* RigidBodyODE.cpp:294-317: dynamic trimesh geom on this body; offset rotation r[0]=M11 r[1]=M21 r[2]=M31 / r[4]=M12 ... (ODE rows =
       mat44f columns), offset position = the matrix' translation.  The reference passes Car::getGraphicsOffsetMatrix() taken while the
       chassis sits at the origin with identity rotation (Car.cpp:342, 1407-1424): GRAPHICS_OFFSET and the GRAPHICS_PITCH_ROTATION rotator. */
    MeshColliderR mc;
    mc.mesh = std::dynamic_pointer_cast<CollisionMeshR>(core->createCollider(trimesh, true, spaceId, category, mask));
    const float r[9] = {offset.M11, offset.M21, offset.M31, offset.M12, offset.M22, offset.M32, offset.M13, offset.M23, offset.M33};
    for (int k = 0; k < 9; ++k) mc.offR[k] = r[k];
    mc.offP[0] = offset.M41; mc.offP[1] = offset.M42; mc.offP[2] = offset.M43;
    meshColliders.push_back(mc);
}

RayCastHit RayCasterR::rayCast(const vec3f& pos, const vec3f& dir) { return core->rayCastImpl(pos, dir, length); }

std::shared_ptr<IPhysicsEngine> PhysicsFactory::createPhysicsEngine() { return std::make_shared<RestateEngine>(); }

/* harness access (ref_harness.cpp) */
oder::World* pdref_world(IPhysicsEngine* e) { return &static_cast<RestateEngine*>(e)->world; }
oder::Body* pdref_body(IRigidBody* rb) { return &static_cast<RigidBodyR*>(rb)->b; }
oder::Joint* pdref_joint(IJoint* j) { return &static_cast<JointR*>(j)->j; }
unsigned int pdref_get_frame(IPhysicsEngine* e) { return static_cast<RestateEngine*>(e)->currentFrame; }
void pdref_set_frame(IPhysicsEngine* e, unsigned int f) { static_cast<RestateEngine*>(e)->currentFrame = f; }
/* contact joints alive after the last step (tests): n x {pos3, normal3, depth, kind}; clearing them = what a state restore implies */
int pdref_engine_contacts(IPhysicsEngine* e, float* out8, int cap) {
    auto* r = static_cast<RestateEngine*>(e); int n = 0;
    for (const oder::ContactJoint& c : r->contactGroupDynamic) { if (n >= cap) break; float* o = out8 + n * 8; o[0] = c.pos[0]; o[1] = c.pos[1]; o[2] = c.pos[2]; o[3] = c.normal[0]; o[4] = c.normal[1]; o[5] = c.normal[2]; o[6] = c.depth; o[7] = c.soft_erp >= 0 ? 0.0f : 1.0f; ++n; }
    return (int)r->contactGroupDynamic.size();
}
void pdref_engine_clear_contacts(IPhysicsEngine* e) { static_cast<RestateEngine*>(e)->contactGroupDynamic.clear(); }
void pdref_engine_set_response(IPhysicsEngine* e, int on) { static_cast<RestateEngine*>(e)->responseEnabled = on != 0; if (!on) static_cast<RestateEngine*>(e)->contactGroupDynamic.clear(); }
/* colliders as the engine received them (for PdCarParams parity) */
int pdref_body_box(IRigidBody* rb, float* centre3, float* size3) {
    auto* r = static_cast<RigidBodyR*>(rb); if (r->boxColliders.empty()) return 0;
    centre3[0] = r->boxColliders[0].centre.x; centre3[1] = r->boxColliders[0].centre.y; centre3[2] = r->boxColliders[0].centre.z;
    size3[0] = r->boxColliders[0].size.x; size3[1] = r->boxColliders[0].size.y; size3[2] = r->boxColliders[0].size.z; return (int)r->boxColliders.size();
}
int pdref_body_mesh(IRigidBody* rb, ITriMesh** mesh, float* offR9, float* off3) {
    auto* r = static_cast<RigidBodyR*>(rb); if (r->meshColliders.empty()) return 0;
    *mesh = r->meshColliders[0].mesh->trimesh.get(); for (int k = 0; k < 3; ++k) off3[k] = r->meshColliders[0].offP[k]; for (int k = 0; k < 9; ++k) offR9[k] = r->meshColliders[0].offR[k]; return (int)r->meshColliders.size();
}
void pdref_ray_stats(IPhysicsEngine* e, unsigned long long* rays, unsigned long long* tris) {
    auto* r = static_cast<RestateEngine*>(e); *rays = r->rayCount; *tris = r->rayTriTests;
}

}
