// See B200MBPolKernels.h.  NOT BUILT here (needs OpenMM); shows exactly which C-ABI call replaces which reference line.
#include "B200MBPolKernels.h"

#include <map>

#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"

using namespace MBPolPlugin;
using namespace OpenMM;

namespace {

void check(int rc, mbpol_ctx* ctx) {
    if (rc != MBPOL_OK) throw OpenMMException(mbpol_last_error(ctx));     // error convention of the reference kernels
}

std::vector<std::vector<int> > particlesOf2B(const MBPolTwoBodyForce& f) {
    std::vector<std::vector<int> > out(f.getNumMolecules());
    for (int i = 0; i < f.getNumMolecules(); ++i) f.getParticleParameters(i, out[i]);   // MBPolReferenceKernels.cpp:345-352
    return out;
}
std::vector<std::vector<int> > particlesOf3B(const MBPolThreeBodyForce& f) {
    std::vector<std::vector<int> > out(f.getNumMolecules());
    for (int i = 0; i < f.getNumMolecules(); ++i) f.getParticleParameters(i, out[i]);
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

void B200SharedEngine::configure2B(const std::vector<std::vector<int> >& particles, double cutoff, int method) {
    indices.clear();
    for (const auto& m : particles) indices.insert(indices.end(), m.begin(), m.begin() + 3);
    cfg.n_molecules = (int)particles.size();
    cfg.cutoff_2b_nm = cutoff;
    cfg.method_2b = method;                  // MBPolTwoBodyForce::NonbondedMethod values == MBPOL_* values
    have2 = true;
}

void B200SharedEngine::configure3B(const std::vector<std::vector<int> >& particles, double cutoff, int method) {
    if (!have2) {
        indices.clear();
        for (const auto& m : particles) indices.insert(indices.end(), m.begin(), m.begin() + 3);
        cfg.n_molecules = (int)particles.size();
    }
    cfg.cutoff_3b_nm = cutoff;
    cfg.method_3b = method;
    have3 = true;
}

void B200SharedEngine::ensureCreated(CudaContext& cu) {
    if (ctx) return;
    cfg.abi_version = MBPOL_ABI_VERSION;
    cfg.n_particles = cu.getNumAtoms();
    cfg.particle_indices = indices.data();
    cfg.device = cu.getDeviceIndex();
    if (!have2) { cfg.cutoff_2b_nm = 1e10; cfg.method_2b = MBPOL_CUTOFF_NONPERIODIC; }
    if (!have3) { cfg.cutoff_3b_nm = 1e10; cfg.method_3b = MBPOL_CUTOFF_NONPERIODIC; }
    check(mbpol_create(&ctx, &cfg), nullptr);
    cudaMalloc(&posBuffer, sizeof(double) * 3 * cu.getNumAtoms());
}

double B200SharedEngine::execute(CudaContext& cu, unsigned flags) {
    ensureCreated(cu);
    cu.setAsCurrent();
    // 1. positions: OpenMM's posq (real4, reordered) -> double3 in original atom order.  A 10-line gather kernel
    //    (posq[atomIndex[i]].xyz -> posBuffer[i], plus posqCorrection in mixed precision) lives in the adapter; the
    //    1e-9 parity bar needs CudaPrecision=double, which is what the reference's CUDA tests use
    //    (platforms/cuda/tests/CMakeLists.txt:20-22).
    gatherPositionsInOriginalOrder(cu, (double*)posBuffer);
    // 2. box: rectangular only, like extractBoxSize (MBPolReferenceKernels.cpp:63-66)
    double4 box = cu.getPeriodicBoxSize();
    double b[3] = {box.x, box.y, box.z};
    if (cfg.method_2b == MBPOL_CUTOFF_PERIODIC || cfg.method_3b == MBPOL_CUTOFF_PERIODIC) check(mbpol_set_box(ctx, b), ctx);
    // 3. evaluate on OpenMM's stream; forces stay inside the library ...
    double e[2];
    check(mbpol_eval(ctx, (const double*)posBuffer, nullptr, e, flags, cu.getCurrentStream()), ctx);
    // 4. ... and are added to OpenMM's int64 fixed-point force buffer (value * 2^32, SoA, padded).  The adapter passes
    //    the atom-order permutation by calling mbpol_update_particles with reordered indices after each reorder.
    check(mbpol_add_forces_fixed_point(ctx, (long long*)cu.getForce().getDevicePointer(), cu.getPaddedNumAtoms(),
                                       cu.getCurrentStream()), ctx);
    return ((flags & MBPOL_DO_2B) ? e[0] : 0.0) + ((flags & MBPOL_DO_3B) ? e[1] : 0.0);
}

void B200SharedEngine::updateParticles(const std::vector<std::vector<int> >& particles) {
    std::vector<int32_t> idx;
    for (const auto& m : particles) idx.insert(idx.end(), m.begin(), m.begin() + 3);
    check(mbpol_update_particles(ctx, (int)particles.size(), idx.data()), ctx);   // MBPOL_ECHANGED -> "number of particles has changed"
}

B200SharedEngine::~B200SharedEngine() {
    if (posBuffer) cudaFree(posBuffer);
    mbpol_destroy(ctx);
}

// ---- CalcMBPolTwoBodyForceKernel ------------------------------------------------------------------------------------
void B200CalcMBPolTwoBodyForceKernel::initialize(const System& system, const MBPolTwoBodyForce& force) {
    numMolecules = force.getNumMolecules();
    std::vector<std::vector<int> > p = particlesOf2B(force);
    B200SharedEngine::get(cu).configure2B(p, force.getCutoff(), (int)force.getNonbondedMethod());
    cu.addForce(new MoleculeGroups(p));
}
double B200CalcMBPolTwoBodyForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    return B200SharedEngine::get(cu).execute(cu, MBPOL_DO_2B);
}
void B200CalcMBPolTwoBodyForceKernel::copyParametersToContext(ContextImpl& context, const MBPolTwoBodyForce& force) {
    if (numMolecules != force.getNumParticles())
        throw OpenMMException("updateParametersInContext: The number of particles has changed");
    B200SharedEngine::get(cu).updateParticles(particlesOf2B(force));
}

// ---- CalcMBPolThreeBodyForceKernel ----------------------------------------------------------------------------------
void B200CalcMBPolThreeBodyForceKernel::initialize(const System& system, const MBPolThreeBodyForce& force) {
    numMolecules = force.getNumMolecules();
    std::vector<std::vector<int> > p = particlesOf3B(force);
    B200SharedEngine::get(cu).configure3B(p, force.getCutoff(), (int)force.getNonbondedMethod());
    cu.addForce(new MoleculeGroups(p));
}
double B200CalcMBPolThreeBodyForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    return B200SharedEngine::get(cu).execute(cu, MBPOL_DO_3B);
}
void B200CalcMBPolThreeBodyForceKernel::copyParametersToContext(ContextImpl& context, const MBPolThreeBodyForce& force) {
    if (numMolecules != force.getNumParticles())
        throw OpenMMException("updateParametersInContext: The number of particles has changed");
    B200SharedEngine::get(cu).updateParticles(particlesOf3B(force));
}

// ---- plugin entry points, same names as platforms/cuda/src/CudaMBPolKernelFactory.cpp:43-69 -------------------------
class B200MBPolKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
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
