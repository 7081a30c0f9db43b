// Tang-Toennies damped C6 dispersion of MB-pol water on the tuple machinery of the 2B path.
//
// The reference does not code this term: python/mbpol.xml:37-108 builds a stock OpenMM CustomNonbondedForce
//     E = sum over atom pairs of different molecules (no M site) with r <= cutoff of   -C6 tt6(r) / r^6
//     tt6(r) = 1 - exp(-d6 r) sum_{k=0..6} (d6 r)^k / k!
// with per-class tables (mbpol.xml:53-82), the cutoff applied to the ATOM-ATOM minimum-image distance, no switching
// function, no long-range correction.  Here: molecule pairs from the same cell-list / neighbour-row builder as the dimer
// list (O-O within cutoff + 0.3 nm, a superset of every molecule pair that has an atom pair within the cutoff), one thread
// per molecule pair evaluating its 9 atom pairs, then the same per-tuple force blocks + fixed-order segmented reduction
// as the 2B term (reduce.cu): deterministic, no floating-point atomics.
#include "context.h"

namespace {

// mbpol.xml:53-82 (kJ/mol nm^6, 1/nm); classes O = 0, H = 1
__constant__ double C6_TAB[2][2] = {{9.92951990e-4, 3.49345451e-4}, {3.49345451e-4, 8.40715638e-5}};
__constant__ double D6_TAB[2][2] = {{9.29548582e+01, 9.77520243e+01}, {9.77520243e+01, 9.40647517e+01}};

__global__ void __launch_bounds__(128) dispersion_pairs_kernel(const double* __restrict__ pos, const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ pairs, int64_t first, int64_t n,
                                                               BoxArg box, double cutoff, double* __restrict__ e_out,
                                                               double* __restrict__ f_out) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int ma = pairs[2 * (first + t)], mb = pairs[2 * (first + t) + 1];
    double a[3][3], b[3][3], fa[3][3] = {}, fb[3][3] = {};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double* pa = pos + 3 * (size_t)idx[3 * ma + k];
        const double* pb = pos + 3 * (size_t)idx[3 * mb + k];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            a[k][c] = pa[c];
            b[k][c] = pb[c];
        }
    }
    double e = 0.0;
    const double c2 = cutoff * cutoff;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double d[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                d[c] = b[j][c] - a[i][c];
                if (box.periodic) d[c] -= box.len[c] * floor(d[c] / box.len[c] + 0.5);    // per atom pair, like OpenMM
            }
            const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (r2 > c2) continue;
            const double c6 = C6_TAB[i > 0][j > 0], d6 = D6_TAB[i > 0][j > 0];
            const double r = sqrt(r2), x = d6 * r;
            // tt6 = 1 - exp(-x) sum_{k<=6} x^k / k!  (Horner);  d tt6 / d r = d6 exp(-x) x^6 / 720
            const double s = 1.0 + x * (1.0 + x * (0.5 + x * (1.0 / 6 + x * (1.0 / 24 + x * (1.0 / 120 + x * (1.0 / 720))))));
            const double ex = exp(-x);
            const double tt = 1.0 - s * ex;
            const double x3 = x * x * x;
            const double dtt = d6 * ex * x3 * x3 * (1.0 / 720);
            const double ir2 = 1.0 / r2, ir6 = ir2 * ir2 * ir2;
            e += -c6 * tt * ir6;
            const double de_over_r = -c6 * (dtt * ir6 - 6.0 * tt * ir6 / r) / r;            // (dE/dr) / r
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                fa[i][c] += de_over_r * d[c];           // F_a = -dE/d(a) = +(dE/dr) d / r with d = b - a
                fb[j][c] -= de_over_r * d[c];
            }
        }
    e_out[t] = e;
    double* f = f_out + 18 * t;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            f[3 * k + c] = fa[k][c];
            f[9 + 3 * k + c] = fb[k][c];
        }
}

}  // namespace

extern "C" int mbpol_eval_dispersion(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, double cutoff_nm,
                                     int32_t periodic, void* stream) {
    if (!ctx || !pos_nm || !(cutoff_nm > 0.0)) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_dispersion: bad argument");
    if (periodic) {
        if (!(ctx->box[0] > 0 && ctx->box[1] > 0 && ctx->box[2] > 0))
            return fail(ctx, MBPOL_EBOX, "mbpol_eval_dispersion: periodic evaluation needs a box (mbpol_set_box)");
        const double m = fmin(ctx->box[0], fmin(ctx->box[1], ctx->box[2]));
        if (cutoff_nm > 0.5 * m)            // OpenMM: "The cutoff distance cannot be greater than half the periodic box size"
            return fail(ctx, MBPOL_EBOX, "CustomNonbondedForce: The cutoff distance cannot be greater than half the periodic box size.");
    }
    cudaStream_t s = (cudaStream_t)stream;
    CKC(cudaSetDevice(ctx->device));
    g_launches = 0;
    NvtxRange range("mbpol:dispersion");
    int rc;
    TupleLists& L = ctx->dlists;
    L.force_two_pass = ctx->lists.force_two_pass;
    const double list_cut = cutoff_nm >= 1e9 ? 1e10 : cutoff_nm + 0.3;       // O-H <= ~0.12 nm on either side, with margin
    if (ctx->n_mol > 0) {
        if ((rc = L.build(pos_nm, ctx->idx.as<int32_t>(), ctx->n_mol, ctx->box, periodic != 0, list_cut, 1e10, true, false, s,
                          ctx->err)))
            return rc;
        if ((rc = L.partition(ctx->rank, ctx->world, s, ctx->err))) return rc;
    }
    const int64_t first = L.own_p0, n = ctx->n_mol > 0 ? L.own_p1 - L.own_p0 : 0;
    if ((rc = ctx->e_d.reserve(sizeof(double) * (size_t)(n > 0 ? n : 1), ctx->err))) return rc;
    if ((rc = ctx->f_d.reserve(sizeof(double) * 18 * (size_t)(n > 0 ? n : 1), ctx->err))) return rc;
    BoxArg box{{ctx->box[0], ctx->box[1], ctx->box[2]}, periodic ? 1 : 0};
    if (n > 0) {
        dispersion_pairs_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(pos_nm, ctx->idx.as<int32_t>(), L.pairs.as<int32_t>(),
                                                                           first, n, box, cutoff_nm, ctx->e_d.as<double>(),
                                                                           ctx->f_d.as<double>());
        MBPOL_COUNT_LAUNCH(1);
    }
    if (ctx->n_mol > 0) {
        if ((rc = ctx->ci_d.build(L.pairs.as<int32_t>(), 2, first, n, ctx->n_mol, L.cub_tmp, s, ctx->err))) return rc;
        if ((rc = ctx->f_mol_d.reserve(sizeof(double) * (9 * (size_t)ctx->n_mol + 9), ctx->err))) return rc;
        launch_force_reduce(ctx->ci_d, 2, ctx->f_d.as<double>(), ctx->n_mol, ctx->f_mol_d.as<double>(), false, s);
        if (force) launch_force_scatter(ctx->f_mol_d.as<double>(), ctx->idx.as<int32_t>(), ctx->n_mol, force, s);
    }
    launch_energy_sum(ctx->e_d.as<double>(), n, ctx->e_partial.as<double>(), ctx->d_energy.as<double>() + 3, s);
    if (energy) {
        CKC(cudaMemcpyAsync(energy, ctx->d_energy.as<double>() + 3, sizeof(double), cudaMemcpyDeviceToHost, s));
        CKC(cudaStreamSynchronize(s));
    }
    ctx->stats.n_launches = g_launches;
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

extern "C" int mbpol_eval_dispersion_host(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, double cutoff_nm,
                                          int32_t periodic) {
    if (!ctx || !pos_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_dispersion_host: null argument");
    CKC(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->n_particles;
    int rc;
    if ((rc = ctx->pos_stage.reserve(bytes + 8, ctx->err))) return rc;
    if ((rc = ctx->force_stage.reserve(bytes + 8, ctx->err))) return rc;
    cudaStream_t s = 0;
    CKC(cudaMemcpyAsync(ctx->pos_stage.ptr, pos_nm, bytes, cudaMemcpyHostToDevice, s));
    if (force) CKC(cudaMemcpyAsync(ctx->force_stage.ptr, force, bytes, cudaMemcpyHostToDevice, s));
    if ((rc = mbpol_eval_dispersion(ctx, ctx->pos_stage.as<double>(), force ? ctx->force_stage.as<double>() : nullptr, energy,
                                    cutoff_nm, periodic, (void*)s)))
        return rc;
    if (force) CKC(cudaMemcpyAsync(force, ctx->force_stage.ptr, bytes, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    return MBPOL_OK;
}
