// stand-in for the plugin's openmmapi/include/openmm/mbpolKernels.h:134-207: the two abstract kernel interfaces
#pragma once
#include <string>
#include "openmm/KernelImpl.h"
#include "openmm/MBPolThreeBodyForce.h"
#include "openmm/MBPolTwoBodyForce.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
namespace MBPolPlugin {
class CalcMBPolTwoBodyForceKernel : public OpenMM::KernelImpl {
public:
    static std::string Name() { return "CalcMBPolTwoBodyForce"; }
    CalcMBPolTwoBodyForceKernel(std::string name, const OpenMM::Platform& platform) : OpenMM::KernelImpl(name, platform) {}
    virtual void initialize(const OpenMM::System& system, const MBPolTwoBodyForce& force) = 0;
    virtual double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy) = 0;
    virtual void copyParametersToContext(OpenMM::ContextImpl& context, const MBPolTwoBodyForce& force) = 0;
};
class CalcMBPolThreeBodyForceKernel : public OpenMM::KernelImpl {
public:
    static std::string Name() { return "CalcMBPolThreeBodyForce"; }
    CalcMBPolThreeBodyForceKernel(std::string name, const OpenMM::Platform& platform) : OpenMM::KernelImpl(name, platform) {}
    virtual void initialize(const OpenMM::System& system, const MBPolThreeBodyForce& force) = 0;
    virtual double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy) = 0;
    virtual void copyParametersToContext(OpenMM::ContextImpl& context, const MBPolThreeBodyForce& force) = 0;
};
}  // namespace MBPolPlugin
