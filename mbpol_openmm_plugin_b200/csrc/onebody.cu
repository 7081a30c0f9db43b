// One-body (intramolecular) term of MB-pol: the Partridge-Schwenke water monomer surface, energy + analytic forces.
//
// Behaviour follows the Reference platform:
//   MBPolReferenceOneBodyForce::calculateOneBodyIxn      platforms/reference/src/MBPolReferenceOneBodyForce.cpp:70-187
//   MBPolReferenceOneBodyForce::calculateForceAndEnergy  platforms/reference/src/MBPolReferenceOneBodyForce.cpp:191-210
//   (periodic: hydrogens imaged to their own oxygen, MBPolReferenceTwoBodyForce.cpp:66-87)
//
// One thread per molecule.  The reference walks three 245-entry exponent tables at run time; here the tables are
// compile-time constants (gen/onebody_terms.inc, written by build.py from data/onebody.npz), so the 3 x 16 power tables
// and their derivatives live in registers and the 244 terms are straight-line FMA code: no local-memory indexing,
// no table loads.  Every particle belongs to one molecule, so forces are added without atomics; per-molecule
// energies go through the same fixed-shape tree sum as the 2B/3B energies (bitwise reproducible).
#include "internal.h"

namespace mbpol {

namespace {

#include "gen/onebody_terms.inc"   // PS_* scalar constants, PS_C0, PS_TERMS(T)

__device__ __forceinline__ double image1(double box, double ref, double p) { return p + box * floor((ref - p) / box + 0.5); }

__global__ void __launch_bounds__(128) onebody_kernel(const double* __restrict__ pos, const int32_t* __restrict__ idx,
                                                      int first, int n, BoxArg box, double* __restrict__ force,
                                                      double* __restrict__ e_mol) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int m = first + t;
    const int pO = __ldg(idx + 3 * m), pH1 = __ldg(idx + 3 * m + 1), pH2 = __ldg(idx + 3 * m + 2);
    double O[3], H1[3], H2[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        O[c] = __ldg(pos + 3 * (size_t)pO + c);
        H1[c] = __ldg(pos + 3 * (size_t)pH1 + c);
        H2[c] = __ldg(pos + 3 * (size_t)pH2 + c);
        if (box.periodic) {
            H1[c] = image1(box.len[c], O[c], H1[c]);
            H2[c] = image1(box.len[c], O[c], H2[c]);
        }
    }
    double r1[3], r2[3], rh[3], d1 = 0, d2 = 0, dh = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        r1[c] = (H1[c] - O[c]) * NM_TO_A;
        r2[c] = (H2[c] - O[c]) * NM_TO_A;
        rh[c] = (H1[c] - H2[c]) * NM_TO_A;
        d1 += r1[c] * r1[c];
        d2 += r2[c] * r2[c];
        dh += rh[c] * rh[c];
    }
    d1 = sqrt(d1);
    d2 = sqrt(d2);
    dh = sqrt(dh);
    const double costh = (r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2]) / (d1 * d2);

    const double deoh = PS_F5Z * PS_DEOHA;
    const double phh1 = PS_F5Z * PS_PHH1A * exp(PS_PHH2);
    const double costhe = -.24780227221366464506;      // cos of the equilibrium angle, MBPolReferenceOneBodyForce.cpp:115
    const double e1 = exp(-PS_ALPHAOH * (d1 - PS_ROH));
    const double e2 = exp(-PS_ALPHAOH * (d2 - PS_ROH));
    const double Va = deoh * (e1 * (e1 - 2.0) + e2 * (e2 - 2.0));
    const double Vb = phh1 * exp(-PS_PHH2 * dh);
    const double dVa1 = 2.0 * PS_ALPHAOH * deoh * e1 * (1.0 - e1) / d1;
    const double dVa2 = 2.0 * PS_ALPHAOH * deoh * e2 * (1.0 - e2) / d2;
    const double dVb = -PS_PHH2 * Vb / dh;

    const double x1 = (d1 - PS_REOH) / PS_REOH, x2 = (d2 - PS_REOH) / PS_REOH, x3 = costh - costhe;
    // power tables: P[0] = 0, P[j] = x^(j-1); D[j] = (j-1) * P[j-1]
    double A[16], B[16], C[16], dA[16], dB[16], dC[16];
    A[0] = B[0] = C[0] = 0.0;
    A[1] = B[1] = C[1] = 1.0;
#pragma unroll
    for (int j = 2; j < 16; ++j) {
        A[j] = A[j - 1] * x1;
        B[j] = B[j - 1] * x2;
        C[j] = C[j - 1] * x3;
    }
    dA[0] = dB[0] = dC[0] = 0.0;
#pragma unroll
    for (int j = 1; j < 16; ++j) {
        dA[j] = (double)(j - 1) * A[j - 1];
        dB[j] = (double)(j - 1) * B[j - 1];
        dC[j] = (double)(j - 1) * C[j - 1];
    }
    const double efac = exp(-PS_B1 * ((d1 - PS_REOH) * (d1 - PS_REOH) + (d2 - PS_REOH) * (d2 - PS_REOH)));
    double sum0 = 0, sum1 = 0, sum2 = 0, sum3 = 0;
#define T(c, i, j, k)                                                   \
    {                                                                   \
        const double ab = A[i] * B[j] + A[j] * B[i];                    \
        sum0 += (c) * ab * C[k];                                        \
        sum1 += (c) * (dA[i] * B[j] + dA[j] * B[i]) * C[k];             \
        sum2 += (c) * (A[i] * dB[j] + A[j] * dB[i]) * C[k];             \
        sum3 += (c) * ab * dC[k];                                       \
    }
    PS_TERMS(T)
#undef T
    const double Vc = 2 * PS_C0 + efac * sum0;
    double e = Va + Vb + Vc;
    e += 0.44739574026257;                    // MBPolReferenceOneBodyForce.cpp:157
    e *= PS_CM1_KCALMOL;
    const double dVcdr1 = (-2 * PS_B1 * efac * (d1 - PS_REOH) * sum0 + efac * sum1 / PS_REOH) / d1;
    const double dVcdr2 = (-2 * PS_B1 * efac * (d2 - PS_REOH) * sum0 + efac * sum2 / PS_REOH) / d2;
    const double dVcdcth = efac * sum3;
    const double unit = PS_CM1_KCALMOL * CAL2JOULE * 10;
    if (force) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double f1 = (dVa1 * r1[c] + dVb * rh[c] + dVcdr1 * r1[c] +
                               dVcdcth * (r2[c] / (d1 * d2) - costh * r1[c] / (d1 * d1))) * unit;
            const double f2 = (dVa2 * r2[c] - dVb * rh[c] + dVcdr2 * r2[c] +
                               dVcdcth * (r1[c] / (d1 * d2) - costh * r2[c] / (d2 * d2))) * unit;
            force[3 * (size_t)pH1 + c] -= f1;
            force[3 * (size_t)pH2 + c] -= f2;
            force[3 * (size_t)pO + c] += f1 + f2;
        }
    }
    e_mol[t] = e * CAL2JOULE;
}

}  // namespace

void launch_onebody(const double* pos, const int32_t* idx, int first, int n, BoxArg box, double* force, double* e_mol,
                    cudaStream_t s) {
    if (n <= 0) return;
    onebody_kernel<<<(n + 127) / 128, 128, 0, s>>>(pos, idx, first, n, box, force, e_mol); MBPOL_COUNT_LAUNCH(1);
}

}  // namespace mbpol
