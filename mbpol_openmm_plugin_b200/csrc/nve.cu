// On-device velocity-Verlet NVE loop around the 2B+3B (+1B) forces (BASELINE config 2 and the NVE
// throughput sweep of config 5; time step and masses as in python/example_ini/mbpol_bulk_nve.ini:11, python/mbpol.xml:3-4):
//
//     v += dt/2 F/m ;  x += dt v ;  M sites placed (average3 weights, python/mbpol.xml:14) ;  F = F(x) ;  v += dt/2 F/m
//
// Positions, velocities, forces and the per-step energies stay in HBM for the whole run: the host only enqueues work
// (the tuple-list build keeps its two 16-byte count read-backs per step) and reads the energy log once at the end.
#include "context.h"

namespace {

// v += (dt/2) F / m, then x += dt v  (massless particles -- virtual sites -- are left alone)
__global__ void kick_drift_kernel(int n3, const double* __restrict__ force, const double* __restrict__ mass, double half_dt,
                                  double dt, double* __restrict__ vel, double* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n3) return;
    const double m = mass[i / 3];
    if (m <= 0.0) return;
    const double v = vel[i] + half_dt * force[i] / m;
    vel[i] = v;
    pos[i] = pos[i] + dt * v;
}

// v += (dt/2) F / m and the particle's kinetic energy (one thread per particle)
__global__ void kick_ke_kernel(int n, const double* __restrict__ force, const double* __restrict__ mass, double half_dt,
                               double* __restrict__ vel, double* __restrict__ ke) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double m = mass[p];
    double e = 0.0;
    if (m > 0.0) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double v = vel[3 * p + c] + half_dt * force[3 * p + c] / m;
            vel[3 * p + c] = v;
            e += m * v * v;
        }
    }
    ke[p] = 0.5 * e;
}

// x_M = w1 x_O + w2 x_H1 + w3 x_H2 (OpenMM ThreeParticleAverageSite)
__global__ void place_msites_kernel(int n_mol, const int32_t* __restrict__ idx, const int32_t* __restrict__ msite, double w1,
                                    double w2, double w3, double* __restrict__ pos) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol || msite[m] < 0) return;
    const double *o = pos + 3 * (size_t)idx[3 * m], *h1 = pos + 3 * (size_t)idx[3 * m + 1], *h2 = pos + 3 * (size_t)idx[3 * m + 2];
    double* x = pos + 3 * (size_t)msite[m];
#pragma unroll
    for (int c = 0; c < 3; ++c) x[c] = w1 * o[c] + w2 * h1[c] + w3 * h2[c];
}

__global__ void log_energy_kernel(const double* __restrict__ e /* E2, E3, E1 */, int with_1b, const double* __restrict__ ke,
                                  double* __restrict__ row) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        row[0] = e[0] + e[1] + (with_1b ? e[2] : 0.0);
        row[1] = *ke;
    }
}

}  // namespace

extern "C" int mbpol_nve_run(mbpol_ctx* ctx, double* pos, double* vel, const double* mass, const mbpol_nve_config* cfg,
                             int32_t n_steps, double* energy_log, void* stream) {
    if (!ctx || !pos || !vel || !mass || !cfg || n_steps < 0) return fail(ctx, MBPOL_EINVAL, "mbpol_nve_run: bad argument");
    const uint32_t many = cfg->flags & (MBPOL_DO_2B | MBPOL_DO_3B);
    const bool with_1b = cfg->flags & MBPOL_NVE_ONEBODY;
    if (!many && !with_1b) return fail(ctx, MBPOL_EINVAL, "mbpol_nve_run: no force term selected");
    if (ctx->world > 1) return fail(ctx, MBPOL_EINVAL, "mbpol_nve_run: single-GPU contexts only");
    cudaStream_t s = (cudaStream_t)stream;
    CKC(cudaSetDevice(ctx->device));
    const int np = ctx->n_particles, n3 = 3 * np;
    int rc;
    if ((rc = ctx->nve_force.reserve(sizeof(double) * n3 + 8, ctx->err))) return rc;
    if ((rc = ctx->nve_ke.reserve(sizeof(double) * (np + 2), ctx->err))) return rc;
    if ((rc = ctx->nve_log.reserve(sizeof(double) * 2 * ((size_t)n_steps + 1), ctx->err))) return rc;
    if (cfg->msite_index) {
        if ((rc = ctx->nve_msite.reserve(sizeof(int32_t) * (size_t)(ctx->n_mol + 1), ctx->err))) return rc;
        CKC(cudaMemcpyAsync(ctx->nve_msite.ptr, cfg->msite_index, sizeof(int32_t) * ctx->n_mol, cudaMemcpyHostToDevice, s));
    }
    double* F = ctx->nve_force.as<double>();
    double* ke_each = ctx->nve_ke.as<double>();
    double* ke_sum = ke_each + np;
    double* log = ctx->nve_log.as<double>();
    const unsigned g3 = (unsigned)((n3 + 255) / 256), gp = (unsigned)((np + 255) / 256), gm = (unsigned)((ctx->n_mol + 255) / 256);
    auto forces = [&]() -> int {
        CKC(cudaMemsetAsync(F, 0, sizeof(double) * n3, s));
        int r;
        if (with_1b && (r = mbpol_eval_onebody(ctx, pos, F, nullptr, cfg->onebody_periodic, s))) return r;
        if (many && (r = mbpol_eval(ctx, pos, F, nullptr, many, s))) return r;      // energy stays on the device
        return MBPOL_OK;
    };
    auto place = [&]() {
        if (cfg->msite_index && ctx->n_mol > 0)
            place_msites_kernel<<<gm, 256, 0, s>>>(ctx->n_mol, ctx->idx.as<int32_t>(), ctx->nve_msite.as<int32_t>(),
                                                   cfg->msite_weights[0], cfg->msite_weights[1], cfg->msite_weights[2], pos);
    };
    const bool timing = ctx->timing;
    ctx->timing = false;                                    // no per-step event synchronisation inside the loop
    NvtxRange range("mbpol:nve");
    place();
    if ((rc = forces())) { ctx->timing = timing; return rc; }
    kick_ke_kernel<<<gp, 256, 0, s>>>(np, F, mass, 0.0, vel, ke_each);           // kinetic energy of the start
    launch_energy_sum(ke_each, np, ctx->e_partial.as<double>(), ke_sum, s);
    log_energy_kernel<<<1, 32, 0, s>>>(ctx->d_energy.as<double>(), with_1b, ke_sum, log);
    for (int step = 1; step <= n_steps; ++step) {
        kick_drift_kernel<<<g3, 256, 0, s>>>(n3, F, mass, 0.5 * cfg->dt_ps, cfg->dt_ps, vel, pos);
        place();
        if ((rc = forces())) { ctx->timing = timing; return rc; }
        kick_ke_kernel<<<gp, 256, 0, s>>>(np, F, mass, 0.5 * cfg->dt_ps, vel, ke_each);
        launch_energy_sum(ke_each, np, ctx->e_partial.as<double>(), ke_sum, s);
        log_energy_kernel<<<1, 32, 0, s>>>(ctx->d_energy.as<double>(), with_1b, ke_sum, log + 2 * (size_t)step);
    }
    ctx->timing = timing;
    if (energy_log) {
        CKC(cudaMemcpyAsync(energy_log, log, sizeof(double) * 2 * ((size_t)n_steps + 1), cudaMemcpyDeviceToHost, s));
        CKC(cudaStreamSynchronize(s));
    }
    CKC(cudaGetLastError());
    return MBPOL_OK;
}
