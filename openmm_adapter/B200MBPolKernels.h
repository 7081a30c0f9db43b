// OpenMM-side adapter: the two kernel classes a CUDA-platform KernelFactory hands out for
//   "CalcMBPolTwoBodyForce" / "CalcMBPolThreeBodyForce"
// implemented on top of the C ABI in include/mbpol_b200.h.  They replace CudaCalcMBPolTwoBodyForceKernel
// (platforms/cuda/src/CudaMBPolKernels.cpp:145-313) and CudaCalcMBPolThreeBodyForceKernel (:1691-2012) of the reference.
//
// Compiled and tested in this repository against the stand-in headers of openmm_adapter/shim (OpenMM is not installed in
// the authoring environment): tests/test_openmm_adapter.py drives these classes through a fake CudaContext
// that owns real device buffers in OpenMM's layouts (shuffled atom order, real4 / double4 posq, padded int64 force
// buffer).  Against a real OpenMM the same two files compile unchanged.
#pragma once

#include <string>
#include <vector>

#include "openmm/mbpolKernels.h"            // the plugin's own header (openmmapi/include/openmm/mbpolKernels.h)
#include "CudaContext.h"                    // OpenMM CUDA platform
#include "mbpol_b200.h"

namespace MBPolPlugin {

// One mbpol_ctx per OpenMM context, shared by the 2B and the 3B kernel: ONE posq gather, ONE tuple-list build and one
// force scatter serve both kernels of a force evaluation (the kernel whose execute() comes first does the work for every
// kernel of the same evaluation whose force group is active; the later one finds its term already added).
class B200SharedEngine {
public:
    static B200SharedEngine& get(OpenMM::CudaContext& cu);
    static void release(OpenMM::CudaContext& cu, unsigned flag);      // a kernel of `cu` is being destroyed
    void configure(unsigned flag, const std::vector<std::vector<int> >& particles, double cutoff, int method, int forceGroup);
    // kernel `flag` (MBPOL_DO_2B | MBPOL_DO_3B) of the evaluation cu.getComputeForceCount(); `groups` = the force groups
    // of that evaluation (ContextImpl::getLastForceGroups()).  Adds forces to cu.getForce() and the energy to
    // cu.getEnergyBuffer() on cu's stream; returns 0.0 like the reference's CUDA kernels (CudaMBPolKernels.cpp:303).
    double execute(OpenMM::CudaContext& cu, unsigned flag, int groups);
    void updateParticles(unsigned flag, const std::vector<std::vector<int> >& particles);
    const mbpol_stats& lastStats() const { return stats; }
    int evaluations() const { return numEvaluations; }
    ~B200SharedEngine();
private:
    void ensureCreated(OpenMM::CudaContext& cu);
    mbpol_ctx* ctx = nullptr;
    mbpol_config cfg{};
    std::vector<int32_t> indices;
    unsigned registered = 0, doneThisEpoch = 0;
    int group2 = 0, group3 = 0, epoch = -1, numEvaluations = 0;
    mbpol_stats stats{};
};

class B200CalcMBPolTwoBodyForceKernel : public CalcMBPolTwoBodyForceKernel {
public:
    B200CalcMBPolTwoBodyForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu,
                                    const OpenMM::System& system)
        : CalcMBPolTwoBodyForceKernel(name, platform), cu(cu), system(system) {}
    ~B200CalcMBPolTwoBodyForceKernel();
    void initialize(const OpenMM::System& system, const MBPolTwoBodyForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const MBPolTwoBodyForce& force);
private:
    OpenMM::CudaContext& cu;
    const OpenMM::System& system;
    int numMolecules = 0;
};

class B200CalcMBPolThreeBodyForceKernel : public CalcMBPolThreeBodyForceKernel {
public:
    B200CalcMBPolThreeBodyForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu,
                                      const OpenMM::System& system)
        : CalcMBPolThreeBodyForceKernel(name, platform), cu(cu), system(system) {}
    ~B200CalcMBPolThreeBodyForceKernel();
    void initialize(const OpenMM::System& system, const MBPolThreeBodyForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const MBPolThreeBodyForce& force);
private:
    OpenMM::CudaContext& cu;
    const OpenMM::System& system;
    int numMolecules = 0;
};

}  // namespace MBPolPlugin

// plugin entry points, same names as platforms/cuda/src/CudaMBPolKernelFactory.cpp:43-69
extern "C" void registerPlatforms();
extern "C" void registerKernelFactories();
extern "C" void registerMBPolCudaKernelFactories();
