// See B200MBPolKernels.h.  Which C-ABI call replaces which reference line is noted at every step.
#include "B200MBPolKernels.h"

#include <map>

#include "CudaPlatform.h"
#include "openmm/KernelFactory.h"
#include "openmm/OpenMMException.h"
#include "openmm/Platform.h"
#include "openmm/internal/ContextImpl.h"

using namespace MBPolPlugin;
using namespace OpenMM;

namespace {

void check(int rc, mbpol_ctx* ctx) {
    if (rc != MBPOL_OK) throw OpenMMException(mbpol_last_error(ctx));     // error convention of the reference kernels
}

template <class Force>
std::vector<std::vector<int> > particlesOf(const Force& f) {
    std::vector<std::vector<int> > out(f.getNumMolecules());
    for (int i = 0; i < f.getNumMolecules(); ++i) f.getParticleParameters(i, out[i]);   // MBPolReferenceKernels.cpp:345-352
    return out;
}

// Keeps molecules intact under OpenMM's atom reordering, like the reference's ForceInfo
// (platforms/cuda/src/CudaMBPolKernels.cpp:147-162, 1693-1711).
class MoleculeGroups : public CudaForceInfo {
public:
    explicit MoleculeGroups(const std::vector<std::vector<int> >& p) : p(p) {}
    int getNumParticleGroups() { return (int)p.size(); }
    void getParticlesInGroup(int index, std::vector<int>& particles) { particles = p[index]; }
    bool areGroupsIdentical(int, int) { return true; }
private:
    std::vector<std::vector<int> > p;
};

std::map<CudaContext*, B200SharedEngine*> engines;

}  // namespace

B200SharedEngine& B200SharedEngine::get(CudaContext& cu) {
    B200SharedEngine*& e = engines[&cu];
    if (!e) e = new B200SharedEngine();
    return *e;
}

// The engine lives as long as one of the kernels of its context does: a CudaContext address that is reused later
// must never find a stale engine.
void B200SharedEngine::release(CudaContext& cu, unsigned flag) {
    std::map<CudaContext*, B200SharedEngine*>::iterator it = engines.find(&cu);
    if (it == engines.end()) return;
    it->second->registered &= ~flag;
    if (!it->second->registered) {
        delete it->second;
        engines.erase(it);
    }
}

void B200SharedEngine::configure(unsigned flag, const std::vector<std::vector<int> >& particles, double cutoff, int method,
                                 int forceGroup) {
    if (ctx) throw OpenMMException("MBPol B200 kernels: initialize() after the first execute()");
    std::vector<int32_t> idx;
    for (size_t m = 0; m < particles.size(); ++m) {
        if (particles[m].size() < 3) throw OpenMMException("MBPol B200 kernels: each molecule needs the particles {O, H1, H2}");
        idx.insert(idx.end(), particles[m].begin(), particles[m].begin() + 3);
    }
    if (registered && idx != indices)
        throw OpenMMException("MBPol B200 kernels: MBPolTwoBodyForce and MBPolThreeBodyForce must list the same molecules");
    indices = idx;
    cfg.n_molecules = (int)particles.size();
    if (flag == MBPOL_DO_2B) {
        cfg.cutoff_2b_nm = cutoff;
        cfg.method_2b = method;                  // MBPolTwoBodyForce::NonbondedMethod values == MBPOL_* values
        group2 = forceGroup;
    } else {
        cfg.cutoff_3b_nm = cutoff;
        cfg.method_3b = method;
        group3 = forceGroup;
    }
    registered |= flag;
}

void B200SharedEngine::ensureCreated(CudaContext& cu) {
    if (ctx) return;
    cfg.abi_version = MBPOL_ABI_VERSION;
    cfg.n_particles = cu.getNumAtoms();
    cfg.particle_indices = indices.data();
    cfg.device = cu.getDeviceIndex();
    if (!(registered & MBPOL_DO_2B)) { cfg.cutoff_2b_nm = 1e10; cfg.method_2b = MBPOL_CUTOFF_NONPERIODIC; }
    if (!(registered & MBPOL_DO_3B)) { cfg.cutoff_3b_nm = 1e10; cfg.method_3b = MBPOL_CUTOFF_NONPERIODIC; }
    check(mbpol_create(&ctx, &cfg), nullptr);
}

double B200SharedEngine::execute(CudaContext& cu, unsigned flag, int groups) {
    ensureCreated(cu);
    cu.setAsCurrent();
    // One pass per force evaluation.  The first of our kernels that OpenMM executes in evaluation number
    // cu.getComputeForceCount() evaluates every registered term whose force group is part of this evaluation; the other
    // kernel's execute() then has nothing left to do.  (Two separate passes would gather posq and build the tuple lists
    // twice per step.)
    const int now = cu.getComputeForceCount();
    if (now != epoch) {
        epoch = now;
        doneThisEpoch = 0;
    }
    if (doneThisEpoch & flag) {
        doneThisEpoch &= ~flag;
        return 0.0;
    }
    unsigned flags = flag;
    const unsigned other = flag == MBPOL_DO_2B ? MBPOL_DO_3B : MBPOL_DO_2B;
    const int otherGroup = other == MBPOL_DO_2B ? group2 : group3;
    if ((registered & other) && (groups & (1 << otherGroup))) flags |= other;
    // box: rectangular only, like extractBoxSize (MBPolReferenceKernels.cpp:63-66)
    if (cfg.method_2b == MBPOL_CUTOFF_PERIODIC || cfg.method_3b == MBPOL_CUTOFF_PERIODIC) {
        const double4 box = cu.getPeriodicBoxSize();
        const double b[3] = {box.x, box.y, box.z};
        check(mbpol_set_box(ctx, b), ctx);
    }
    // posq (reordered, real4 / double4) -> original-order double3 -> lists + 2B + 3B -> int64 fixed-point force buffer and
    // energy buffer, all on OpenMM's stream, nothing copied to the host (CudaMBPolKernels.cpp:278-304, 1976-2005)
    check(mbpol_eval_openmm(ctx, cu.getPosq().getDevicePointer(), cu.getUseDoublePrecision() ? 1 : 0,
                            (const int32_t*)cu.getAtomIndexArray().getDevicePointer(), cu.getNumAtoms(),
                            (long long*)cu.getForce().getDevicePointer(), cu.getPaddedNumAtoms(),
                            cu.getEnergyBuffer().getDevicePointer(),
                            (cu.getUseDoublePrecision() || cu.getUseMixedPrecision()) ? 1 : 0, flags,
                            (void*)cu.getCurrentStream()),
          ctx);
    mbpol_get_stats(ctx, &stats);
    ++numEvaluations;
    doneThisEpoch = flags & ~flag;               // terms evaluated on behalf of the other kernel
    return 0.0;
}

void B200SharedEngine::updateParticles(unsigned flag, const std::vector<std::vector<int> >& particles) {
    std::vector<int32_t> idx;
    for (size_t m = 0; m < particles.size(); ++m) idx.insert(idx.end(), particles[m].begin(), particles[m].begin() + 3);
    indices = idx;
    if (ctx) check(mbpol_update_particles(ctx, (int)particles.size(), idx.data()), ctx);   // MBPOL_ECHANGED -> "number of particles has changed"
}

B200SharedEngine::~B200SharedEngine() { mbpol_destroy(ctx); }

// ---- CalcMBPolTwoBodyForceKernel ------------------------------------------------------------------------------------
B200CalcMBPolTwoBodyForceKernel::~B200CalcMBPolTwoBodyForceKernel() { B200SharedEngine::release(cu, MBPOL_DO_2B); }
void B200CalcMBPolTwoBodyForceKernel::initialize(const System& system, const MBPolTwoBodyForce& force) {
    numMolecules = force.getNumMolecules();
    std::vector<std::vector<int> > p = particlesOf(force);
    B200SharedEngine::get(cu).configure(MBPOL_DO_2B, p, force.getCutoff(), (int)force.getNonbondedMethod(), force.getForceGroup());
    cu.addForce(new MoleculeGroups(p));
}
double B200CalcMBPolTwoBodyForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    return B200SharedEngine::get(cu).execute(cu, MBPOL_DO_2B, context.getLastForceGroups());
}
void B200CalcMBPolTwoBodyForceKernel::copyParametersToContext(ContextImpl& context, const MBPolTwoBodyForce& force) {
    if (numMolecules != force.getNumParticles())
        throw OpenMMException("updateParametersInContext: The number of particles has changed");
    B200SharedEngine::get(cu).updateParticles(MBPOL_DO_2B, particlesOf(force));
}

// ---- CalcMBPolThreeBodyForceKernel ----------------------------------------------------------------------------------
B200CalcMBPolThreeBodyForceKernel::~B200CalcMBPolThreeBodyForceKernel() { B200SharedEngine::release(cu, MBPOL_DO_3B); }
void B200CalcMBPolThreeBodyForceKernel::initialize(const System& system, const MBPolThreeBodyForce& force) {
    numMolecules = force.getNumMolecules();
    std::vector<std::vector<int> > p = particlesOf(force);
    B200SharedEngine::get(cu).configure(MBPOL_DO_3B, p, force.getCutoff(), (int)force.getNonbondedMethod(), force.getForceGroup());
    cu.addForce(new MoleculeGroups(p));
}
double B200CalcMBPolThreeBodyForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    return B200SharedEngine::get(cu).execute(cu, MBPOL_DO_3B, context.getLastForceGroups());
}
void B200CalcMBPolThreeBodyForceKernel::copyParametersToContext(ContextImpl& context, const MBPolThreeBodyForce& force) {
    if (numMolecules != force.getNumParticles())
        throw OpenMMException("updateParametersInContext: The number of particles has changed");
    B200SharedEngine::get(cu).updateParticles(MBPOL_DO_3B, particlesOf(force));
}

// ---- plugin entry points, same names as platforms/cuda/src/CudaMBPolKernelFactory.cpp:43-69 -------------------------
class B200MBPolKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
        // like the reference, contexts[0] drives the whole evaluation (CudaMBPolKernelFactory.cpp:72); several GPUs are
        // used INSIDE the library (mbpol_comm_*), not through OpenMM's per-context force split
        CudaContext& cu = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData())->contexts[0];
        if (name == CalcMBPolTwoBodyForceKernel::Name())
            return new B200CalcMBPolTwoBodyForceKernel(name, platform, cu, context.getSystem());
        if (name == CalcMBPolThreeBodyForceKernel::Name())
            return new B200CalcMBPolThreeBodyForceKernel(name, platform, cu, context.getSystem());
        throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '") + name + "'").c_str());
    }
};

extern "C" void registerPlatforms() {}
extern "C" void registerKernelFactories() {
    try {
        Platform& platform = Platform::getPlatformByName("CUDA");
        B200MBPolKernelFactory* factory = new B200MBPolKernelFactory();
        platform.registerKernelFactory(CalcMBPolTwoBodyForceKernel::Name(), factory);
        platform.registerKernelFactory(CalcMBPolThreeBodyForceKernel::Name(), factory);
    } catch (...) {
    }
}
extern "C" void registerMBPolCudaKernelFactories() { registerKernelFactories(); }
