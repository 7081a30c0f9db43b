#pragma once
#include "openmm/System.h"
namespace OpenMM {
class ContextImpl {
public:
    ContextImpl(const System& system, void* platformData) : system(system), platformData(platformData), lastForceGroups(-1) {}
    const System& getSystem() const { return system; }
    void* getPlatformData() { return platformData; }
    int getLastForceGroups() const { return lastForceGroups; }
    void setLastForceGroups(int groups) { lastForceGroups = groups; }       // (ContextImpl::calcForcesAndEnergy does this)
private:
    const System& system;
    void* platformData;
    int lastForceGroups;
};
}  // namespace OpenMM
