// Internal declarations shared by the translation units of libmbpol_b200.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/mbpol_b200.h"

namespace mbpol {

// kernels launched since the counter was last reset (mbpol_stats::n_launches); every <<<>>> site bumps it
extern thread_local int64_t g_launches;
#define MBPOL_COUNT_LAUNCH(n) (::mbpol::g_launches += (n))

// ---- physical constants of the path --------------------------------------------------------------------------------
// Unit conversions: openmmapi/include/openmm/internal/MBPolConstants.h:37-38
constexpr double NM_TO_A = 10.0;
constexpr double CAL2JOULE = 4.184;

// 2B nonlinear parameters: platforms/reference/src/mbpol_2body_constants.h:34-49 (values are physical parameters of
// the published MB-pol model, Babin/Leforestier/Paesani JCTC 2013)
constexpr double K2_HH_INTRA = -6.480884773303821e-01;
constexpr double K2_OH_INTRA = 1.674518993682975e+00;
constexpr double K2_HH_COUL = 1.148231864355956e+00;
constexpr double K2_OH_COUL = 1.205989761123099e+00;
constexpr double K2_OO_COUL = 1.395357065790959e+00;
constexpr double K2_XH_MAIN = 7.347036852042255e-01;
constexpr double K2_XO_MAIN = 7.998249864422826e-01;
constexpr double K2_XX_MAIN = 7.960663960630585e-01;
constexpr double GAMMA_IN_PLANE = -9.721486914088159e-02;
constexpr double GAMMA_OUT_OF_PLANE = 9.859272078406150e-02;
constexpr double R2I = 4.5, R2F = 6.5;          // 2B switch range, Angstrom
constexpr double D0_INTRA = 1.0, D0_INTER = 4.0;

// 3B nonlinear parameters: platforms/reference/src/mbpol_3body_constants.h:5-20 (Babin/Medders/Paesani JCTC 2014)
constexpr double R3I = 0.0, R3F = 4.5;
constexpr double K3_HH_INTRA = -2.254303257872797e+00, D3_HH_INTRA = 1.690594510379166e+00;
constexpr double K3_OH_INTRA = -2.564771901404151e-01, D3_OH_INTRA = -2.999999868517452e+00;
constexpr double K3_HH = 4.247920544074333e-01, D3_HH = 3.499031358429095e+00;
constexpr double K3_OH = 8.128985941165371e-01, D3_OH = 4.854042963514281e+00;
constexpr double K3_OO = 3.749457984616480e-02, D3_OO = 4.816312044947604e-08;

constexpr int NV2 = 31, NV3 = 36;

// ---- tuple stage (tuples.cu) ---------------------------------------------------------------------------------------
struct BoxArg {
    double len[3];   // nm
    int periodic;
};

// variables x[nv][ld] (tuple index fastest)
void launch_dimer_vars(const double* pos, const int32_t* idx, const int32_t* pairs, int64_t first, int64_t n, int64_t ld,
                       BoxArg box, double* x, cudaStream_t s);
void launch_trimer_vars(const double* pos, const int32_t* idx, const int32_t* trips, int64_t first, int64_t n, int64_t ld,
                        BoxArg box, double* x, cudaStream_t s);
// chain rule + switch from poly[1+nv][ld] = {E, x_i dE/dx_i}: per-tuple energies e[n] and force blocks f[n][18|27]
void launch_dimer_forces(const double* pos, const int32_t* idx, const int32_t* pairs, int64_t first, int64_t n, int64_t ld,
                         BoxArg box, const double* poly, double* e, double* f, cudaStream_t s);
void launch_trimer_forces(const double* pos, const int32_t* idx, const int32_t* trips, int64_t first, int64_t n,
                          int64_t ld, BoxArg box, const double* poly, double* e, double* f, cudaStream_t s);

// ---- one-body term (onebody.cu): molecules [first, first+n), forces added in place, e_mol[n] in kJ/mol ----------------
void launch_onebody(const double* pos, const int32_t* idx, int first, int n, BoxArg box, double* force, double* e_mol,
                    cudaStream_t s);

// ---- polynomial stage (pip_module.cu): generated PTX kernels loaded through the driver API --------------------------
// out[0][t] = E, out[1 + i][t] = dE/dx_i, or x_i dE/dx_i with scale_by_x (what the force kernels consume)
int pip_launch(int order, const double* x, double* out, int64_t n, int64_t ld, cudaStream_t s, std::string* err,
               bool scale_by_x = false);
int pip_instruction_count(int order);

}  // namespace mbpol
