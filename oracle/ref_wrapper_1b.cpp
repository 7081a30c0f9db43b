/* TEST INFRASTRUCTURE ONLY -- C-ABI wrapper around the UNMODIFIED reference one-body source
 * (MBPolReferenceOneBodyForce.cpp:191-210 calculateForceAndEnergy, :70-187 calculateOneBodyIxn; constants
 * mbpol_interaction_constants.h:22-424), built by oracle/Makefile into oracle/_ref/. Never linked into the product. */
#include "MBPolReferenceOneBodyForce.h"
#include <vector>

using namespace OpenMM;

extern "C" double ref_eval_1b(const double* pos, double* forces, long n_particles, const int* idx, int n_mol,
                              const double* box, int pbc) {
    std::vector<RealVec> p(n_particles), f(n_particles);
    for (long i = 0; i < n_particles; ++i) {
        p[i] = RealVec(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
        f[i] = RealVec(forces[3 * i], forces[3 * i + 1], forces[3 * i + 2]);
    }
    std::vector<std::vector<int> > ids(n_mol, std::vector<int>(3));
    for (int m = 0; m < n_mol; ++m) for (int k = 0; k < 3; ++k) ids[m][k] = idx[3 * m + k];
    MBPolReferenceOneBodyForce force;
    if (pbc) {
        force.setNonbondedMethod(MBPolReferenceOneBodyForce::Periodic);
        force.setPeriodicBox(RealVec(box[0], box[1], box[2]));
    }
    double e = force.calculateForceAndEnergy(n_mol, p, ids, f);
    for (long i = 0; i < n_particles; ++i) for (int k = 0; k < 3; ++k) forces[3 * i + k] = f[i][k];
    return e;
}
