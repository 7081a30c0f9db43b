/* TEST INFRASTRUCTURE ONLY -- C-ABI wrapper around the UNMODIFIED reference sources (compiled where they lie
 * under /root/reference by oracle/Makefile into oracle/_ref/libmbpol_ref.so). Never linked into the product.
 *
 * This TU textually includes the reference's ReferenceThreeNeighborList.cpp so that, besides the exported
 * computeThreeNeighborListVoxelHash (ReferenceThreeNeighborList.cpp:181-226), the file-local ThreeVoxelHash
 * (:60-177) can be driven directly to obtain the *pair* relation nearby[i] = {j<i : d2 <= c2} with the plugin's
 * own copy of OpenMM's voxel-hash algorithm (OpenMM's computeNeighborListVoxelHash itself is not in the tree,
 * SURVEY.md section 8c).  It also wraps MBPolReferenceThreeBodyForce (MBPolReferenceThreeBodyForce.cpp:295-322)
 * and poly_3b_v2x::eval (poly-3b-v2x.cpp:3). */
#include "ReferenceThreeNeighborList.cpp"
#include "MBPolReferenceThreeBodyForce.h"
#include "mbpol_3body_constants.h"
#include "poly-3b-v2x.h"
#include <cstring>

using namespace MBPolPlugin;
using namespace OpenMM;

static std::vector<RealVec> to_vec(const double* p, long n) {
    std::vector<RealVec> v(n);
    for (long i = 0; i < n; ++i) v[i] = RealVec(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
    return v;
}

extern "C" {

/* triplets in the reference's native emission order; returns the count, writes at most cap triplets */
long ref_triplets(const double* posO, int n, const double* box, int pbc, double cutoff, unsigned* out, long cap) {
    std::vector<RealVec> loc = to_vec(posO, n);
    RealVec b(box[0], box[1], box[2]);
    ThreeNeighborList list;
    computeThreeNeighborListVoxelHash(list, n, loc, b, pbc != 0, cutoff, 0.0);
    long m = (long)list.size();
    for (long t = 0; t < m && t < cap; ++t) {
        out[3 * t] = list[t].first; out[3 * t + 1] = list[t].second; out[3 * t + 2] = list[t].third;
    }
    return m;
}

/* pairs (i, j<i) produced by the reference's ThreeVoxelHash exactly as computeThreeNeighborListVoxelHash drives it
 * (ReferenceThreeNeighborList.cpp:190-213): native order, (first=i higher index, second=j) */
long ref_pairs(const double* posO, int n, const double* box, int pbc, double cutoff, unsigned* out, long cap) {
    std::vector<RealVec> loc = to_vec(posO, n);
    RealVec b(box[0], box[1], box[2]);
    std::vector<std::vector<AtomIndex> > nearby(n);
    double ex, ey, ez;
    if (!pbc) ex = ey = ez = cutoff;
    else {
        ex = b[0] / floor(b[0] / cutoff); ey = b[1] / floor(b[1] / cutoff); ez = b[2] / floor(b[2] / cutoff);
    }
    ThreeVoxelHash hash(ex, ey, ez, b, pbc != 0);
    for (AtomIndex j = 0; j < (AtomIndex)n; ++j) {
        hash.getNeighbors(nearby, VoxelItem(&loc[j], j), cutoff, 0.0);
        hash.insert(j, loc[j]);
    }
    long m = 0;
    for (int i = 0; i < n; ++i)
        for (size_t k = 0; k < nearby[i].size(); ++k, ++m)
            if (m < cap) { out[2 * m] = i; out[2 * m + 1] = nearby[i][k]; }
    return m;
}

/* 3B energy (kJ/mol) and forces (+=, kJ/mol/nm) over a given triplet list.
 * idx = n_mol x 3 global particle indices {O,H1,H2}; pos/forces = n_particles x 3 */
double ref_eval_3b(const double* pos, double* forces, long n_particles, const int* idx, int n_mol,
                   const unsigned* trip, long n_trip, const double* box, int pbc, double cutoff) {
    std::vector<RealVec> p = to_vec(pos, n_particles), f = to_vec(forces, n_particles);
    std::vector<std::vector<int> > ids(n_mol, std::vector<int>(3));
    for (int m = 0; m < n_mol; ++m) for (int k = 0; k < 3; ++k) ids[m][k] = idx[3 * m + k];
    ThreeNeighborList list(n_trip);
    for (long t = 0; t < n_trip; ++t) { list[t].first = trip[3 * t]; list[t].second = trip[3 * t + 1]; list[t].third = trip[3 * t + 2]; }
    MBPolReferenceThreeBodyForce force;
    force.setCutoff(cutoff);
    if (pbc) {
        force.setNonbondedMethod(MBPolReferenceThreeBodyForce::CutoffPeriodic);
        force.setPeriodicBox(RealVec(box[0], box[1], box[2]));
    } else force.setNonbondedMethod(MBPolReferenceThreeBodyForce::CutoffNonPeriodic);
    double e = force.calculateForceAndEnergy(n_mol, p, ids, list, f);
    for (long i = 0; i < n_particles; ++i) for (int k = 0; k < 3; ++k) forces[3 * i + k] = f[i][k];
    return e;
}

/* per-triplet energies and 9x3 force blocks (slots Oa,Ha1,Ha2,Ob,Hb1,Hb2,Oc,Hc1,Hc2), particles = stride-3 molecules */
void ref_each_3b(const double* pos, int n_mol, const unsigned* trip, long n_trip, const double* box, int pbc,
                 double* e_each, double* f_each) {
    std::vector<RealVec> p = to_vec(pos, 3L * n_mol);
    std::vector<std::vector<int> > ids(n_mol, std::vector<int>(3));
    for (int m = 0; m < n_mol; ++m) for (int k = 0; k < 3; ++k) ids[m][k] = 3 * m + k;
    MBPolReferenceThreeBodyForce force;
    if (pbc) {
        force.setNonbondedMethod(MBPolReferenceThreeBodyForce::CutoffPeriodic);
        force.setPeriodicBox(RealVec(box[0], box[1], box[2]));
    } else force.setNonbondedMethod(MBPolReferenceThreeBodyForce::CutoffNonPeriodic);
    std::vector<RealVec> f(3L * n_mol);
    ThreeNeighborList one(1);
    for (long t = 0; t < n_trip; ++t) {
        one[0].first = trip[3 * t]; one[0].second = trip[3 * t + 1]; one[0].third = trip[3 * t + 2];
        unsigned mol[3] = {one[0].first, one[0].second, one[0].third};
        for (int s = 0; s < 3; ++s) for (int k = 0; k < 3; ++k) f[3 * mol[s] + k] = RealVec();
        e_each[t] = force.calculateForceAndEnergy(n_mol, p, ids, one, f);
        for (int s = 0; s < 3; ++s) for (int k = 0; k < 3; ++k) for (int c = 0; c < 3; ++c)
            f_each[27 * t + 9 * s + 3 * k + c] = f[3 * mol[s] + k][c];
    }
}

double ref_poly3b(const double* x, double* g) { return poly_3b_v2x::eval(thefit, x, g); }
int ref_ncoef3b(void) { return (int)(sizeof(thefit) / sizeof(double)); }
const double* ref_coef3b(void) { return thefit; }

}  // extern "C"
