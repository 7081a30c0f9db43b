// stand-in for the plugin's openmmapi/include/openmm/MBPolTwoBodyForce.h:57-166 (same methods, same defaults)
#pragma once
#include <vector>
namespace MBPolPlugin {
class MBPolTwoBodyForce {
public:
    enum NonbondedMethod { NoCutoff = 0, CutoffPeriodic = 1, CutoffNonPeriodic = 2 };
    MBPolTwoBodyForce() : nonbondedMethod(CutoffNonPeriodic), cutoff(1.0e+10), forceGroup(0) {}
    int getNumParticles() const { return (int)parameters.size(); }
    int getNumMolecules() const { return (int)parameters.size(); }
    int addParticle(const std::vector<int>& particleIndices) { parameters.push_back(particleIndices); return (int)parameters.size() - 1; }
    void getParticleParameters(int particleIndex, std::vector<int>& particleIndices) const { particleIndices = parameters[particleIndex]; }
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
}  // namespace MBPolPlugin
