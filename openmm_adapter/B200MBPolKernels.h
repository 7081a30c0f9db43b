// OpenMM-side adapter: the two kernel classes a CUDA-platform KernelFactory hands out for
//   "CalcMBPolTwoBodyForce" / "CalcMBPolThreeBodyForce"
// implemented on top of the C ABI in include/mbpol_b200.h.  They replace CudaCalcMBPolTwoBodyForceKernel
// (platforms/cuda/src/CudaMBPolKernels.cpp:145-313) and CudaCalcMBPolThreeBodyForceKernel (:1691-2012) of the reference.
//
// NOT BUILT in this repository's CI: OpenMM headers/libraries are not available in the authoring container
// (SURVEY.md 8c).  Everything OpenMM-independent that these classes do is exercised through the same C ABI by
// mbpol_openmm_plugin_b200/api.py (same method names) and tests/.  See INTEGRATION.md for the build lines.
#pragma once

#include <string>
#include <vector>

#include "openmm/mbpolKernels.h"            // reference header, unchanged (openmmapi/include/openmm/mbpolKernels.h)
#include "CudaContext.h"                    // OpenMM CUDA platform
#include "mbpol_b200.h"

namespace MBPolPlugin {

// One mbpol_ctx per OpenMM context, shared by the 2B and the 3B kernel so that one list build serves both.
class B200SharedEngine {
public:
    static B200SharedEngine& get(OpenMM::CudaContext& cu);
    void configure2B(const std::vector<std::vector<int> >& particles, double cutoff, int method);
    void configure3B(const std::vector<std::vector<int> >& particles, double cutoff, int method);
    // evaluates `flags` (MBPOL_DO_2B | MBPOL_DO_3B) and adds the forces into cu.getForce(); returns the energy
    double execute(OpenMM::CudaContext& cu, unsigned flags);
    void updateParticles(const std::vector<std::vector<int> >& particles);
    ~B200SharedEngine();
private:
    void ensureCreated(OpenMM::CudaContext& cu);
    mbpol_ctx* ctx = nullptr;
    mbpol_config cfg{};
    std::vector<int32_t> indices;
    bool have2 = false, have3 = false;
    void* posBuffer = nullptr;               // n_particles x 3 doubles in ORIGINAL atom order (see .cpp)
};

class B200CalcMBPolTwoBodyForceKernel : public CalcMBPolTwoBodyForceKernel {
public:
    B200CalcMBPolTwoBodyForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu,
                                    const OpenMM::System& system)
        : CalcMBPolTwoBodyForceKernel(name, platform), cu(cu), system(system) {}
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
    void initialize(const OpenMM::System& system, const MBPolThreeBodyForce& force);
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    void copyParametersToContext(OpenMM::ContextImpl& context, const MBPolThreeBodyForce& force);
private:
    OpenMM::CudaContext& cu;
    const OpenMM::System& system;
    int numMolecules = 0;
};

}  // namespace MBPolPlugin
