#pragma once
#include <vector>
namespace OpenMM {
class System {
public:
    int addParticle(double mass) { masses.push_back(mass); return (int)masses.size() - 1; }
    int getNumParticles() const { return (int)masses.size(); }
    double getParticleMass(int i) const { return masses[i]; }
private:
    std::vector<double> masses;
};
}  // namespace OpenMM
