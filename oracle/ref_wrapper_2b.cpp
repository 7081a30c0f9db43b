/* TEST INFRASTRUCTURE ONLY -- C-ABI wrapper around the UNMODIFIED reference 2B sources
 * (MBPolReferenceTwoBodyForce.cpp:298-323 calculateForceAndEnergy, poly-2b-v6x.cpp:3 poly_2b_v6x_eval,
 * mbpol_2body_constants.h:54-1208 thefit), built by oracle/Makefile into oracle/_ref/. Never linked into the product. */
#include "MBPolReferenceTwoBodyForce.h"
#include "mbpol_2body_constants.h"
#include "poly-2b-v6x.h"
#include <vector>

using namespace OpenMM;

static std::vector<RealVec> to_vec(const double* p, long n) {
    std::vector<RealVec> v(n);
    for (long i = 0; i < n; ++i) v[i] = RealVec(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}

static void configure(MBPolReferenceTwoBodyForce& force, const double* box, int pbc) {
    if (pbc) {
        force.setNonbondedMethod(MBPolReferenceTwoBodyForce::CutoffPeriodic);
        force.setPeriodicBox(RealVec(box[0], box[1], box[2]));
    } else force.setNonbondedMethod(MBPolReferenceTwoBodyForce::CutoffNonPeriodic);
}

extern "C" {

double ref_eval_2b(const double* pos, double* forces, long n_particles, const int* idx, int n_mol,
                   const unsigned* pairs, long n_pairs, const double* box, int pbc, double cutoff) {
    std::vector<RealVec> p = to_vec(pos, n_particles), f = to_vec(forces, n_particles);
    std::vector<std::vector<int> > ids(n_mol, std::vector<int>(3));
    for (int m = 0; m < n_mol; ++m) for (int k = 0; k < 3; ++k) ids[m][k] = idx[3 * m + k];
    NeighborList list(n_pairs);
    for (long t = 0; t < n_pairs; ++t) list[t] = AtomPair(pairs[2 * t], pairs[2 * t + 1]);
    MBPolReferenceTwoBodyForce force;
    force.setCutoff(cutoff);
    configure(force, box, pbc);
    double e = force.calculateForceAndEnergy(n_mol, p, ids, list, f);
    for (long i = 0; i < n_particles; ++i) for (int k = 0; k < 3; ++k) forces[3 * i + k] = f[i][k];
    return e;
}

/* per-pair energies and 6x3 force blocks (slots Oa,Ha1,Ha2,Ob,Hb1,Hb2), particles = stride-3 molecules */
void ref_each_2b(const double* pos, int n_mol, const unsigned* pairs, long n_pairs, const double* box, int pbc,
                 double* e_each, double* f_each) {
    std::vector<RealVec> p = to_vec(pos, 3L * n_mol);
    std::vector<std::vector<int> > ids(n_mol, std::vector<int>(3));
    for (int m = 0; m < n_mol; ++m) for (int k = 0; k < 3; ++k) ids[m][k] = 3 * m + k;
    MBPolReferenceTwoBodyForce force;
    configure(force, box, pbc);
    std::vector<RealVec> f(3L * n_mol);
    NeighborList one(1);
    for (long t = 0; t < n_pairs; ++t) {
        one[0] = AtomPair(pairs[2 * t], pairs[2 * t + 1]);
        unsigned mol[2] = {one[0].first, one[0].second};
        for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) f[3 * mol[s] + k] = RealVec();
        e_each[t] = force.calculateForceAndEnergy(n_mol, p, ids, one, f);
        for (int s = 0; s < 2; ++s) for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c)
            f_each[18 * t + 9 * s + 3 * k + c] = f[3 * mol[s] + k][c];
    }
}

double ref_poly2b(const double* x, double* g) { return poly_2b_v6x_eval(thefit, x, g); }
int ref_ncoef2b(void) { return (int)(sizeof(thefit) / sizeof(double)); }
const double* ref_coef2b(void) { return thefit; }

}  // extern "C"
