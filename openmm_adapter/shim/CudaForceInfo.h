#pragma once
#include <vector>
namespace OpenMM {
class CudaForceInfo {
public:
    virtual ~CudaForceInfo() {}
    virtual bool areParticlesIdentical(int, int) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int, std::vector<int>&) {}
    virtual bool areGroupsIdentical(int, int) { return true; }
};
}  // namespace OpenMM
