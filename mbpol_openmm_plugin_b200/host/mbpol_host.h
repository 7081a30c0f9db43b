// C++ host layer above the C ABI (include/mbpol_b200.h), mirroring the plugin's operator interface for the 2B+3B path
// without depending on an OpenMM install (OpenMM is absent in this environment):
//
//   MBPolPlugin::MBPolTwoBodyForce / MBPolThreeBodyForce   same methods and defaults as
//       openmmapi/include/openmm/MBPolTwoBodyForce.h:57-166, MBPolThreeBodyForce.h:57-166
//   MBPolPlugin::B200CalcMBPolTwoBodyForceKernel / ...ThreeBody...   Name() / initialize / execute /
//       copyParametersToContext of openmmapi/include/openmm/mbpolKernels.h:134-207
//   MBPolPlugin::B200Context    the slice of OpenMM::Context the path needs (positions, periodic box vectors,
//       getState(Forces | Energy)), including the ForceImpl checks of openmmapi/src/MBPolTwoBodyForceImpl.cpp:55-76
//   OpenMM::Vec3, OpenMM::OpenMMException   minimal stand-ins with the real names, so that tests written against
//       this header read like platforms/reference/tests/TestReferenceMBPol{Two,Three}BodyForce.cpp
//
// With a real OpenMM the same kernels are registered through openmm_adapter/B200MBPolKernels.cpp.
#pragma once

#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "mbpol_b200.h"

namespace OpenMM {

class OpenMMException : public std::runtime_error {
public:
    explicit OpenMMException(const std::string& msg) : std::runtime_error(msg) {}
};

class Vec3 {
public:
    Vec3() : v{0, 0, 0} {}
    Vec3(double x, double y, double z) : v{x, y, z} {}
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    Vec3 operator+(const Vec3& o) const { return Vec3(v[0] + o[0], v[1] + o[1], v[2] + o[2]); }
    Vec3 operator-(const Vec3& o) const { return Vec3(v[0] - o[0], v[1] - o[1], v[2] - o[2]); }
    Vec3 operator*(double s) const { return Vec3(v[0] * s, v[1] * s, v[2] * s); }
    Vec3& operator/=(double s) { v[0] /= s; v[1] /= s; v[2] /= s; return *this; }
    Vec3& operator*=(double s) { v[0] *= s; v[1] *= s; v[2] *= s; return *this; }
private:
    double v[3];
};

}  // namespace OpenMM

namespace MBPolPlugin {

// parameter container shared by the two forces (openmmapi/src/MBPolTwoBodyForce.cpp:43-81)
class MBPolManyBodyForceBase {
public:
    enum NonbondedMethod { NoCutoff = 0, CutoffPeriodic = 1, CutoffNonPeriodic = 2 };
    MBPolManyBodyForceBase() : nonbondedMethod(CutoffNonPeriodic), cutoff(1.0e+10), forceGroup(0) {}
    int getNumParticles() const { return (int)parameters.size(); }
    int getNumMolecules() const { return (int)parameters.size(); }
    int addParticle(const std::vector<int>& particleIndices) {
        parameters.push_back(particleIndices);
        return (int)parameters.size() - 1;
    }
    void getParticleParameters(int particleIndex, std::vector<int>& particleIndices) const {
        particleIndices = parameters[particleIndex];
    }
    void setParticleParameters(int particleIndex, std::vector<int>& particleIndices) { parameters[particleIndex] = particleIndices; }
    void setCutoff(double c) { cutoff = c; }
    double getCutoff() const { return cutoff; }
    NonbondedMethod getNonbondedMethod() const { return nonbondedMethod; }
    void setNonbondedMethod(NonbondedMethod m) { nonbondedMethod = m; }
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int g) { forceGroup = g; }
private:
    NonbondedMethod nonbondedMethod;
    double cutoff;
    int forceGroup;
    std::vector<std::vector<int> > parameters;
};
class MBPolTwoBodyForce : public MBPolManyBodyForceBase {};
class MBPolThreeBodyForce : public MBPolManyBodyForceBase {};

// openmmapi/include/openmm/MBPolOneBodyForce.h: one {O,H1,H2} entry per molecule
class MBPolOneBodyForce {
public:
    enum NonbondedMethod { NonPeriodic = 0, Periodic = 1 };
    MBPolOneBodyForce() : nonbondedMethod(NonPeriodic), forceGroup(0) {}
    int getNumOneBodys() const { return (int)stretchBends.size(); }
    int addOneBody(const std::vector<int>& particleIndices) {
        stretchBends.push_back(particleIndices);
        return (int)stretchBends.size() - 1;
    }
    void getOneBodyParameters(int index, std::vector<int>& particleIndices) const { particleIndices = stretchBends[index]; }
    void setOneBodyParameters(int index, std::vector<int>& particleIndices) { stretchBends[index] = particleIndices; }
    NonbondedMethod getNonbondedMethod() const { return nonbondedMethod; }
    void setNonbondedMethod(NonbondedMethod m) { nonbondedMethod = m; }
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int g) { forceGroup = g; }
private:
    NonbondedMethod nonbondedMethod;
    int forceGroup;
    std::vector<std::vector<int> > stretchBends;
};

class B200Context;

// one mbpol_ctx shared by the two kernels of a context
class B200Engine {
public:
    ~B200Engine() { mbpol_destroy(ctx); }
    void check(int rc) const {
        if (rc != MBPOL_OK) throw OpenMM::OpenMMException(mbpol_last_error(ctx));
    }
    mbpol_ctx* ctx = nullptr;
    mbpol_config cfg{};
    std::vector<int32_t> indices;
    bool created = false;
};

class B200KernelBase {
public:
    B200KernelBase(std::string name, B200Context& context, unsigned flag) : name(name), context(context), flag(flag) {}
    virtual ~B200KernelBase() {}
    const std::string& getName() const { return name; }
    double execute(B200Context& context, bool includeForces, bool includeEnergy);
    int forceGroup = 0;
protected:
    void initializeCommon(const MBPolManyBodyForceBase& force, bool twoBody);
    void copyCommon(const MBPolManyBodyForceBase& force);
    std::string name;
    B200Context& context;
    unsigned flag;
    int numMolecules = 0;
};

class B200CalcMBPolTwoBodyForceKernel : public B200KernelBase {
public:
    static std::string Name() { return "CalcMBPolTwoBodyForce"; }
    explicit B200CalcMBPolTwoBodyForceKernel(B200Context& c) : B200KernelBase(Name(), c, MBPOL_DO_2B) {}
    void initialize(const MBPolTwoBodyForce& force) { initializeCommon(force, true); }
    void copyParametersToContext(B200Context&, const MBPolTwoBodyForce& force) { copyCommon(force); }
};
class B200CalcMBPolThreeBodyForceKernel : public B200KernelBase {
public:
    static std::string Name() { return "CalcMBPolThreeBodyForce"; }
    explicit B200CalcMBPolThreeBodyForceKernel(B200Context& c) : B200KernelBase(Name(), c, MBPOL_DO_3B) {}
    void initialize(const MBPolThreeBodyForce& force) { initializeCommon(force, false); }
    void copyParametersToContext(B200Context&, const MBPolThreeBodyForce& force) { copyCommon(force); }
};

// CalcMBPolOneBodyForceKernel (openmmapi/include/openmm/mbpolKernels.h:47-82; MBPolReferenceKernels.cpp:77-126)
class B200CalcMBPolOneBodyForceKernel {
public:
    static std::string Name() { return "CalcMBPolOneBodyForce"; }
    explicit B200CalcMBPolOneBodyForceKernel(B200Context& c) : context(c) {}
    void initialize(const MBPolOneBodyForce& force);
    void copyParametersToContext(B200Context&, const MBPolOneBodyForce& force);
    int forceGroup = 0;
    bool usePBC = false;
    int numOneBodys = 0;
private:
    static std::vector<int32_t> gather(const MBPolOneBodyForce& force);
    B200Context& context;
};

struct B200State {
    double potentialEnergy = 0.0;
    std::vector<OpenMM::Vec3> forces;
    double getPotentialEnergy() const { return potentialEnergy; }
    const std::vector<OpenMM::Vec3>& getForces() const { return forces; }
};

class B200Context {
public:
    // numParticles = System::getNumParticles(); box vectors like System::setDefaultPeriodicBoxVectors (rectangular)
    B200Context(int numParticles, const MBPolTwoBodyForce* f2, const MBPolThreeBodyForce* f3, const OpenMM::Vec3* boxVectors = nullptr,
                const MBPolOneBodyForce* f1 = nullptr)
        : numParticles(numParticles), engine(new B200Engine) {
        if (boxVectors) { hasBox = true; box[0] = boxVectors[0][0]; box[1] = boxVectors[1][1]; box[2] = boxVectors[2][2]; }
        engine->cfg.abi_version = MBPOL_ABI_VERSION;
        engine->cfg.n_particles = numParticles;
        engine->cfg.device = -1;
        engine->cfg.cutoff_2b_nm = engine->cfg.cutoff_3b_nm = 1e10;
        engine->cfg.method_2b = engine->cfg.method_3b = MBPOL_CUTOFF_NONPERIODIC;
        // MBPol*ForceImpl::initialize: periodic cutoff check, then createKernel(Name()) + kernel.initialize(system, owner)
        if (f2) { checkCutoff(*f2); k2.reset(new B200CalcMBPolTwoBodyForceKernel(*this)); k2->forceGroup = f2->getForceGroup(); k2->initialize(*f2); }
        if (f3) { checkCutoff(*f3); k3.reset(new B200CalcMBPolThreeBodyForceKernel(*this)); k3->forceGroup = f3->getForceGroup(); k3->initialize(*f3); }
        if (f1) {
            if (f1->getNonbondedMethod() == MBPolOneBodyForce::Periodic && !hasBox)
                throw OpenMM::OpenMMException("MBPolOneBodyForce: Periodic needs periodic box vectors");
            k1.reset(new B200CalcMBPolOneBodyForceKernel(*this)); k1->forceGroup = f1->getForceGroup(); k1->initialize(*f1);
        }
        if (!f1 && !f2 && !f3) throw OpenMM::OpenMMException("no MBPol force in the system");
        engine->cfg.particle_indices = engine->indices.data();
        int rc = mbpol_create(&engine->ctx, &engine->cfg);
        if (rc != MBPOL_OK) throw OpenMM::OpenMMException(mbpol_last_error(nullptr));
        engine->created = true;
        if (hasBox) engine->check(mbpol_set_box(engine->ctx, box));
    }
    void setPositions(const std::vector<OpenMM::Vec3>& p) {
        if ((int)p.size() != numParticles) throw OpenMM::OpenMMException("Called setPositions() with the wrong number of positions");
        positions = p;
    }
    void setPeriodicBoxVectors(const OpenMM::Vec3& a, const OpenMM::Vec3& b, const OpenMM::Vec3& c) {
        hasBox = true; box[0] = a[0]; box[1] = b[1]; box[2] = c[2];
        engine->check(mbpol_set_box(engine->ctx, box));
    }
    // Context::getState(State::Forces | State::Energy, false, groups)
    B200State getState(int groups = 0xFFFFFFFF) {
        unsigned flags = 0;
        if (k2 && (groups & (1 << k2->forceGroup))) flags |= MBPOL_DO_2B;
        if (k3 && (groups & (1 << k3->forceGroup))) flags |= MBPOL_DO_3B;
        B200State st;
        st.forces.assign(numParticles, OpenMM::Vec3());
        static_assert(sizeof(OpenMM::Vec3) == 3 * sizeof(double), "Vec3 must be three packed doubles");
        if (k1 && (groups & (1 << k1->forceGroup))) {
            double e1 = 0;
            engine->check(mbpol_eval_onebody_host(engine->ctx, &positions[0][0], &st.forces[0][0], &e1, k1->usePBC ? 1 : 0));
            st.potentialEnergy += e1;
        }
        if (!flags) return st;
        double e[2] = {0, 0};
        engine->check(mbpol_eval_host(engine->ctx, &positions[0][0], &st.forces[0][0], e, flags));
        st.potentialEnergy += e[0] + e[1];
        return st;
    }
    B200Engine& getEngine() { return *engine; }
private:
    void checkCutoff(const MBPolManyBodyForceBase& f) const {
        if (f.getNonbondedMethod() != MBPolManyBodyForceBase::CutoffPeriodic) return;
        if (!hasBox) throw OpenMM::OpenMMException("MBPolForce: CutoffPeriodic needs periodic box vectors");
        double c = f.getCutoff();          // openmmapi/src/MBPolTwoBodyForceImpl.cpp:60-66
        if (c > 0.5 * box[0] || c > 0.5 * box[1] || c > 0.5 * box[2])
            throw OpenMM::OpenMMException("MBPolTwoBodyForce: The cutoff distance cannot be greater than half the periodic box size.");
    }
    int numParticles;
    bool hasBox = false;
    double box[3] = {0, 0, 0};
    std::vector<OpenMM::Vec3> positions;
    std::unique_ptr<B200Engine> engine;
    std::unique_ptr<B200CalcMBPolOneBodyForceKernel> k1;
    std::unique_ptr<B200CalcMBPolTwoBodyForceKernel> k2;
    std::unique_ptr<B200CalcMBPolThreeBodyForceKernel> k3;
    friend class B200KernelBase;
    friend class B200CalcMBPolOneBodyForceKernel;
};

inline std::vector<int32_t> B200CalcMBPolOneBodyForceKernel::gather(const MBPolOneBodyForce& force) {
    std::vector<int32_t> idx;
    for (int i = 0; i < force.getNumOneBodys(); ++i) {
        std::vector<int> p;
        force.getOneBodyParameters(i, p);
        if (p.size() < 3) throw OpenMM::OpenMMException("MBPolOneBodyForce: each entry needs particle indices {O, H1, H2}");
        idx.insert(idx.end(), p.begin(), p.begin() + 3);
    }
    return idx;
}

inline void B200CalcMBPolOneBodyForceKernel::initialize(const MBPolOneBodyForce& force) {
    // ReferenceCalcMBPolOneBodyForceKernel::initialize (MBPolReferenceKernels.cpp:86-99)
    B200Engine& e = *context.engine;
    numOneBodys = force.getNumOneBodys();
    std::vector<int32_t> idx = gather(force);
    if (!e.indices.empty() && e.indices != idx) throw OpenMM::OpenMMException("MBPolForce: the forces must list the same molecules");
    e.indices = idx;
    e.cfg.n_molecules = numOneBodys;
    usePBC = force.getNonbondedMethod() == MBPolOneBodyForce::Periodic;
}

inline void B200CalcMBPolOneBodyForceKernel::copyParametersToContext(B200Context&, const MBPolOneBodyForce& force) {
    if (numOneBodys != force.getNumOneBodys())       // MBPolReferenceKernels.cpp:117-118
        throw OpenMM::OpenMMException("updateParametersInContext: The number of stretch-bends has changed");
    B200Engine& e = *context.engine;
    std::vector<int32_t> idx = gather(force);
    e.check(mbpol_update_particles(e.ctx, numOneBodys, idx.data()));
    e.indices = idx;
}

inline void B200KernelBase::initializeCommon(const MBPolManyBodyForceBase& force, bool twoBody) {
    // ReferenceCalcMBPol*ForceKernel::initialize (platforms/reference/src/MBPolReferenceKernels.cpp:341-360, 430-449)
    B200Engine& e = *context.engine;
    numMolecules = force.getNumMolecules();
    std::vector<int32_t> idx;
    for (int i = 0; i < numMolecules; ++i) {
        std::vector<int> p;
        force.getParticleParameters(i, p);
        if (p.size() < 3) throw OpenMM::OpenMMException("MBPolForce: each molecule needs particle indices {O, H1, H2}");
        idx.insert(idx.end(), p.begin(), p.begin() + 3);
    }
    if (!e.indices.empty() && e.indices != idx) throw OpenMM::OpenMMException("MBPolForce: the two forces must list the same molecules");
    e.indices = idx;
    e.cfg.n_molecules = numMolecules;
    if (twoBody) { e.cfg.cutoff_2b_nm = force.getCutoff(); e.cfg.method_2b = (int)force.getNonbondedMethod(); }
    else { e.cfg.cutoff_3b_nm = force.getCutoff(); e.cfg.method_3b = (int)force.getNonbondedMethod(); }
}

inline void B200KernelBase::copyCommon(const MBPolManyBodyForceBase& force) {
    if (numMolecules != force.getNumParticles())
        throw OpenMM::OpenMMException("updateParametersInContext: The number of particles has changed");
    B200Engine& e = *context.engine;
    std::vector<int32_t> idx;
    for (int i = 0; i < numMolecules; ++i) {
        std::vector<int> p;
        force.getParticleParameters(i, p);
        idx.insert(idx.end(), p.begin(), p.begin() + 3);
    }
    e.check(mbpol_update_particles(e.ctx, numMolecules, idx.data()));
    e.indices = idx;
}

inline double B200KernelBase::execute(B200Context& ctx, bool, bool) {
    // like the Reference platform, forces and energy are always both evaluated (MBPolReferenceKernels.cpp:362-399)
    B200State st = ctx.getState(1 << forceGroup);
    return st.potentialEnergy;
}

}  // namespace MBPolPlugin
