// C ABI of libmbpol_b200.so (include/mbpol_b200.h): context management and the per-step pipeline
//
//   gather O -> cell list -> neighbour rows -> pairs / chain triplets      (lists.cu)
//   per tuple: variables -> PIP + gradient (generated sm_100a PTX) -> chain rule + switch   (tuples.cu, pip_module.cu)
//   per molecule: fixed-order segmented force reduction, fixed-shape energy sum              (reduce.cu)
//
// which is what ReferenceCalcMBPol{Two,Three}BodyForceKernel::execute does on the CPU
// (platforms/reference/src/MBPolReferenceKernels.cpp:362-399, 451-483).
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "context.h"

thread_local std::string g_create_error;

namespace {

inline int64_t round_up(int64_t n, int64_t m) { return (n + m - 1) / m * m; }

bool is_periodic(int method) { return method == MBPOL_CUTOFF_PERIODIC; }

// Every particle may belong to one molecule only: the force scatter and the one-body kernel write force[p] without
// atomics (the reference accumulates serially, so it has no such constraint to state).  Returns an error text or "".
std::string check_indices(const int32_t* idx, int64_t n_mol, int64_t n_particles) {
    std::vector<char> seen((size_t)n_particles, 0);
    for (int64_t i = 0; i < 3 * n_mol; ++i) {
        const int32_t p = idx[i];
        if (p < 0 || p >= n_particles) return "particle index out of range";
        if (seen[p]) return "particle index " + std::to_string(p) + " listed more than once";
        seen[p] = 1;
    }
    return "";
}
double eff_cutoff(int method, double c) { return method == MBPOL_NOCUTOFF ? 1e10 : c; }

// One body order over the owned tuple range.  order = 2 | 3.  part: ALL, or the chain split in two so that two orders can
// be enqueued on two streams -- COMPUTE (variables, polynomial, per-tuple force blocks, contribution index) and REDUCE
// (segmented force reduction into f_mol, energy sum).  side: the chain uses the second set of scratch buffers.
enum { PART_ALL = 0, PART_COMPUTE = 1, PART_REDUCE = 2 };
int run_order(mbpol_ctx* ctx, int order, const double* pos, cudaStream_t s, bool first_into_fmol, int part = PART_ALL,
              bool side = false) {
    TupleLists& L = ctx->lists;
    const int width = order;
    const int nv = order == 2 ? NV2 : NV3;
    const int64_t first = order == 2 ? L.own_p0 : L.own_t0;
    const int64_t n = order == 2 ? L.own_p1 - L.own_p0 : L.own_t1 - L.own_t0;
    const int32_t* tuples = order == 2 ? L.pairs.as<int32_t>() : L.trips.as<int32_t>();
    DeviceBuf &X = order == 2 ? ctx->x2 : ctx->x3, &O = order == 2 ? ctx->o2 : ctx->o3,
              &E = order == 2 ? ctx->e2 : ctx->e3, &F = order == 2 ? ctx->f2 : ctx->f3;
    ContributionIndex& ci = order == 2 ? ctx->ci2 : ctx->ci3;
    const int64_t ld = round_up(n > 0 ? n : 1, 32);
    int rc;
    std::string& err = ctx->err;
    if ((rc = X.reserve(sizeof(double) * nv * ld, err))) return rc;
    if ((rc = O.reserve(sizeof(double) * (nv + 1) * ld, err))) return rc;
    if ((rc = E.reserve(sizeof(double) * ld, err))) return rc;
    if ((rc = F.reserve(sizeof(double) * 9 * width * ld, err))) return rc;
    BoxArg box;
    for (int d = 0; d < 3; ++d) box.len[d] = ctx->box[d];
    box.periodic = is_periodic(order == 2 ? ctx->method2 : ctx->method3);
    const int32_t* idx = ctx->idx.as<int32_t>();
    if (part == PART_REDUCE) goto reduce;
    if (order == 2) {
        launch_dimer_vars(pos, idx, tuples, first, n, ld, box, X.as<double>(), s);
        if (ctx->timing) cudaEventRecord(ctx->evp[0], s);
        if ((rc = pip_launch(2, X.as<double>(), O.as<double>(), n, ld, s, &err, true))) return rc;
        if (ctx->timing) cudaEventRecord(ctx->evp[1], s);
        launch_dimer_forces(pos, idx, tuples, first, n, ld, box, O.as<double>(), E.as<double>(), F.as<double>(), s);
    } else {
        launch_trimer_vars(pos, idx, tuples, first, n, ld, box, X.as<double>(), s);
        if (ctx->timing) cudaEventRecord(ctx->evp[2], s);
        if ((rc = pip_launch(3, X.as<double>(), O.as<double>(), n, ld, s, &err, true))) return rc;
        if (ctx->timing) cudaEventRecord(ctx->evp[3], s);
        launch_trimer_forces(pos, idx, tuples, first, n, ld, box, O.as<double>(), E.as<double>(), F.as<double>(), s);
    }
    if ((rc = ci.build(tuples, width, first, n, ctx->n_mol, side ? ctx->cub_tmp2 : L.cub_tmp, s, err))) return rc;
    if (part == PART_COMPUTE) goto done;
reduce:
    launch_force_reduce(ci, width, F.as<double>(), ctx->n_mol, ctx->f_mol.as<double>(), !first_into_fmol, s);
    launch_energy_sum(E.as<double>(), n, (side ? ctx->e_partial2 : ctx->e_partial).as<double>(),
                      ctx->d_energy.as<double>() + (order - 2), s);
done:
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return fail(ctx, MBPOL_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
    }
    return MBPOL_OK;
}

int check_box(mbpol_ctx* ctx, double cutoff) {
    // MBPolReferenceKernels.cpp:384-391 / 469-475
    const double min_allowed = 1.999999 * cutoff;
    if (ctx->box[0] < min_allowed || ctx->box[1] < min_allowed || ctx->box[2] < min_allowed)
        return fail(ctx, MBPOL_EBOX, "The periodic box size has decreased to less than twice the cutoff.");
    return MBPOL_OK;
}

int eval_pass(mbpol_ctx* ctx, const double* pos, bool want2, bool want3, bool periodic, uint32_t flags, cudaStream_t s,
              bool& fmol_started) {
    int rc;
    const double c2 = eff_cutoff(ctx->method2, ctx->cut2), c3 = eff_cutoff(ctx->method3, ctx->cut3);
    if (periodic) {
        if (want2 && (rc = check_box(ctx, c2))) return rc;
        if (want3 && (rc = check_box(ctx, c3))) return rc;
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[0], s);
    NvtxRange r_lists("mbpol:lists");
    const bool reuse = (flags & MBPOL_KEEP_LISTS) && ctx->lists_valid && ctx->lists.has2 == want2 &&
                       ctx->lists.has3 == want3;
    if (!reuse) {
        ctx->lists_valid = false;
        // While the balance of the last repartition is valid, every build is requested SLICED.  All ranks take this
        // decision from the same counters, so their owner ranges always tile [0, n_mol).
        const bool slice = ctx->world > 1 && ctx->slice_valid && ctx->slice_want2 == want2 && ctx->slice_want3 == want3 &&
                           ctx->slice_periodic == periodic && ctx->slice_age < ctx->repartition_every;
        if ((rc = ctx->lists.build(pos, ctx->idx.as<int32_t>(), ctx->n_mol, ctx->box, periodic, c2, c3, want2, want3, s,
                                   ctx->err, slice ? ctx->slice_lo : 0, slice ? ctx->slice_hi : -1)))
            return rc;
        if (slice) {
            // A requested slice may still come back as full lists on THIS rank only (a neighbour row overflowed the
            // fixed stride, or the range is all of [0, n)).  The owner range and the refresh schedule must not move:
            // the other ranks keep theirs.  Only the offsets of the range inside the full lists are looked up.
            if (!ctx->lists.sliced &&
                (rc = ctx->lists.partition(ctx->rank, ctx->world, s, ctx->err, ctx->slice_lo, ctx->slice_hi)))
                return rc;
            ++ctx->slice_age;
        } else {
            if ((rc = ctx->lists.partition(ctx->rank, ctx->world, s, ctx->err))) return rc;
            ctx->slice_valid = ctx->world > 1 && ctx->repartition_every > 0;
            ctx->slice_lo = ctx->lists.own_i0;
            ctx->slice_hi = ctx->lists.own_i1;
            ctx->slice_want2 = want2;
            ctx->slice_want3 = want3;
            ctx->slice_periodic = periodic;
            ctx->slice_age = 0;
            ++ctx->n_repartitions;
        }
        ctx->lists_valid = true;
    }
    r_lists.end();
    if (ctx->timing) cudaEventRecord(ctx->ev[1], s);
    if (want2 && want3 && !ctx->timing && !ctx->serial_orders && ctx->side) {
        // Two chains, two streams.  Issue order = priority of the block scheduler: the 3B chain first (its polynomial
        // kernel takes every SM, wave after wave), the 2B chain on `side` fills the SMs that the last, partial 3B wave
        // leaves idle, and the small kernels of either chain run next to the other's polynomial kernel.  f_mol keeps its
        // order of additions (2B written first, 3B added): the 3B reduction waits for the 2B chain.
        NvtxRange r("mbpol:2B||3B");
        const bool first = !fmol_started;
        if (cudaEventRecord(ctx->ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0) != cudaSuccess)
            return fail(ctx, MBPOL_ECUDA, "two-stream fork failed");
        if ((rc = run_order(ctx, 3, pos, s, false, PART_COMPUTE))) return rc;
        if ((rc = run_order(ctx, 2, pos, ctx->side, first, PART_ALL, true))) return rc;
        if (cudaEventRecord(ctx->ev_join, ctx->side) != cudaSuccess || cudaStreamWaitEvent(s, ctx->ev_join, 0) != cudaSuccess)
            return fail(ctx, MBPOL_ECUDA, "two-stream join failed");
        if ((rc = run_order(ctx, 3, pos, s, false, PART_REDUCE))) return rc;
        fmol_started = true;
        want2 = want3 = false;                               // (the counters below still need the flags: restored there)
        ctx->stats.n_pairs = ctx->lists.sliced ? -1 : ctx->lists.n_pairs;
        ctx->stats.n_pairs_owned = ctx->lists.own_p1 - ctx->lists.own_p0;
        ctx->stats.n_triplets = ctx->lists.sliced ? -1 : ctx->lists.n_trips;
        ctx->stats.n_triplets_owned = ctx->lists.own_t1 - ctx->lists.own_t0;
    }
    if (want2) {
        NvtxRange r("mbpol:2B");
        if ((rc = run_order(ctx, 2, pos, s, !fmol_started))) return rc;
        fmol_started = true;
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[2], s);
    if (want3) {
        NvtxRange r("mbpol:3B");
        if ((rc = run_order(ctx, 3, pos, s, !fmol_started))) return rc;
        fmol_started = true;
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[3], s);
    // after a sliced build only this rank's tuples exist: the box-wide totals are then unknown (-1), not stale
    if (want2) {
        ctx->stats.n_pairs = ctx->lists.sliced ? -1 : ctx->lists.n_pairs;
        ctx->stats.n_pairs_owned = ctx->lists.own_p1 - ctx->lists.own_p0;
    }
    if (want3) {
        ctx->stats.n_triplets = ctx->lists.sliced ? -1 : ctx->lists.n_trips;
        ctx->stats.n_triplets_owned = ctx->lists.own_t1 - ctx->lists.own_t0;
    }
    return MBPOL_OK;
}

}  // namespace

int mbpol_collect_timing(mbpol_ctx* ctx) {
    CKC(cudaEventSynchronize(ctx->ev[4]));
    cudaEventElapsedTime(&ctx->stats.ms_lists, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&ctx->stats.ms_2b, ctx->ev[1], ctx->ev[2]);
    cudaEventElapsedTime(&ctx->stats.ms_3b, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&ctx->stats.ms_reduce, ctx->ev[3], ctx->ev[4]);
    if (ctx->timed2) cudaEventElapsedTime(&ctx->stats.ms_pip_2b, ctx->evp[0], ctx->evp[1]);
    if (ctx->timed3) cudaEventElapsedTime(&ctx->stats.ms_pip_3b, ctx->evp[2], ctx->evp[3]);
    return MBPOL_OK;
}

extern "C" {

int mbpol_create(mbpol_ctx** out, const mbpol_config* cfg) {
    if (!out || !cfg) return fail(nullptr, MBPOL_EINVAL, "mbpol_create: null argument");
    *out = nullptr;
    if (cfg->abi_version != MBPOL_ABI_VERSION) return fail(nullptr, MBPOL_EINVAL, "mbpol_create: ABI version mismatch");
    if (cfg->n_molecules < 0 || cfg->n_particles < 0 || (cfg->n_molecules > 0 && !cfg->particle_indices))
        return fail(nullptr, MBPOL_EINVAL, "mbpol_create: bad molecule / particle counts");
    for (int m : {cfg->method_2b, cfg->method_3b})
        if (m < 0 || m > 2) return fail(nullptr, MBPOL_EINVAL, "mbpol_create: unknown nonbonded method");
    if (!(cfg->cutoff_2b_nm > 0) || !(cfg->cutoff_3b_nm > 0))
        return fail(nullptr, MBPOL_EINVAL, "mbpol_create: cutoffs must be positive");
    {
        const std::string bad = check_indices(cfg->particle_indices, cfg->n_molecules, cfg->n_particles);
        if (!bad.empty()) return fail(nullptr, MBPOL_EINVAL, "mbpol_create: " + bad);
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, MBPOL_ECUDA, "mbpol_create: no CUDA device (this library has no CPU fallback)");
    }
    mbpol_ctx* ctx = new mbpol_ctx;
    if (cfg->device >= 0) ctx->device = cfg->device; else cudaGetDevice(&ctx->device);
    if (ctx->device >= ndev) {
        delete ctx;
        return fail(nullptr, MBPOL_EINVAL, "mbpol_create: device ordinal out of range");
    }
    ctx->n_mol = cfg->n_molecules;
    ctx->n_particles = cfg->n_particles;
    ctx->cut2 = cfg->cutoff_2b_nm;
    ctx->cut3 = cfg->cutoff_3b_nm;
    ctx->method2 = cfg->method_2b;
    ctx->method3 = cfg->method_3b;
    auto bail = [&](int code, const std::string& msg) {
        delete ctx;
        return fail(nullptr, code, msg);
    };
    if (cudaSetDevice(ctx->device) != cudaSuccess) return bail(MBPOL_ECUDA, "cudaSetDevice failed");
    std::string err;
    if (ctx->idx.reserve(sizeof(int32_t) * (3 * (size_t)ctx->n_mol + 4), err) ||
        ctx->f_mol.reserve(sizeof(double) * (9 * (size_t)ctx->n_mol + 9), err) ||
        ctx->e_partial.reserve(sizeof(double) * 1024, err) || ctx->d_energy.reserve(sizeof(double) * 8, err))
        return bail(MBPOL_ENOMEM, err);
    if (ctx->n_mol > 0 && cudaMemcpy(ctx->idx.ptr, cfg->particle_indices, sizeof(int32_t) * 3 * (size_t)ctx->n_mol,
                                     cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(MBPOL_ECUDA, "copying particle indices failed");
    cudaMemset(ctx->d_energy.ptr, 0, 4 * sizeof(double));
    if (const char* f = getenv("MBPOL_FORCE_TWO_PASS")) ctx->lists.force_two_pass = f[0] == '1';
    if (const char* f = getenv("MBPOL_REPARTITION_EVERY")) ctx->repartition_every = atoi(f);   // 0: never slice
    if (const char* f = getenv("MBPOL_SERIAL_ORDERS")) ctx->serial_orders = f[0] == '1';
    if (ctx->e_partial2.reserve(sizeof(double) * 1024, err)) return bail(MBPOL_ENOMEM, err);
    if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess)
        return bail(MBPOL_ECUDA, "creating the side stream failed");
    for (auto& e : ctx->ev) cudaEventCreate(&e);
    for (auto& e : ctx->evp) cudaEventCreate(&e);
    for (auto& e : ctx->evc) cudaEventCreate(&e);
    *out = ctx;
    return MBPOL_OK;
}

void mbpol_destroy(mbpol_ctx* ctx) {
    if (!ctx) return;
    mbpol_comm_destroy(ctx);
    cudaSetDevice(ctx->device);
    for (auto* set : {ctx->ev + 0, ctx->evp + 0, ctx->evc + 0})
        for (int i = 0; i < (set == ctx->ev ? 6 : 4); ++i)
            if (set[i]) cudaEventDestroy(set[i]);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->side2) cudaStreamDestroy(ctx->side2);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (auto& e : ctx->ev_slice)
        if (e) cudaEventDestroy(e);
    delete ctx;
}

const char* mbpol_last_error(const mbpol_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int mbpol_update_particles(mbpol_ctx* ctx, int32_t n_molecules, const int32_t* particle_indices) {
    if (!ctx || !particle_indices) return fail(ctx, MBPOL_EINVAL, "mbpol_update_particles: null argument");
    if (n_molecules != ctx->n_mol)
        return fail(ctx, MBPOL_ECHANGED, "updateParametersInContext: The number of particles has changed");
    {
        const std::string bad = check_indices(particle_indices, n_molecules, ctx->n_particles);
        if (!bad.empty()) return fail(ctx, MBPOL_EINVAL, "mbpol_update_particles: " + bad);
    }
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaMemcpy(ctx->idx.ptr, particle_indices, sizeof(int32_t) * 3 * (size_t)n_molecules, cudaMemcpyHostToDevice));
    ctx->lists_valid = false;
    ctx->slice_valid = false;
    return MBPOL_OK;
}

int mbpol_set_box(mbpol_ctx* ctx, const double box_nm[3]) {
    if (!ctx || !box_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_set_box: null argument");
    for (int d = 0; d < 3; ++d) {
        if (!(box_nm[d] > 0) || !std::isfinite(box_nm[d])) return fail(ctx, MBPOL_EINVAL, "mbpol_set_box: box edges must be positive");
        ctx->box[d] = box_nm[d];
    }
    ctx->lists_valid = false;
    return MBPOL_OK;
}

int mbpol_set_skin(mbpol_ctx* ctx, double skin_nm) {
    if (!ctx || !(skin_nm >= 0.0) || !std::isfinite(skin_nm)) return fail(ctx, MBPOL_EINVAL, "mbpol_set_skin: bad argument");
    ctx->lists.skin = skin_nm;
    ctx->lists.skin_valid = false;
    return MBPOL_OK;
}

int mbpol_set_partition(mbpol_ctx* ctx, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || rank < 0 || rank >= world) return fail(ctx, MBPOL_EINVAL, "mbpol_set_partition: bad rank/world");
    ctx->rank = rank;
    ctx->world = world;
    ctx->lists_valid = false;
    ctx->slice_valid = false;
    return MBPOL_OK;
}

int mbpol_eval(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, uint32_t flags, void* stream) {
    if (!ctx || !pos_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval: null argument");
    const bool want2 = flags & MBPOL_DO_2B, want3 = flags & MBPOL_DO_3B;
    if (!want2 && !want3) return fail(ctx, MBPOL_EINVAL, "mbpol_eval: neither MBPOL_DO_2B nor MBPOL_DO_3B requested");
    cudaStream_t s = (cudaStream_t)stream;
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaMemsetAsync(ctx->d_energy.ptr, 0, 2 * sizeof(double), s));
    ctx->stats = mbpol_stats{};
    ctx->f_mol_valid = false;
    g_launches = 0;
    bool started = false;
    int rc;
    const bool p2 = is_periodic(ctx->method2), p3 = is_periodic(ctx->method3);
    if (ctx->n_mol > 0) {
        if (want2 && want3 && p2 != p3) {       // the two forces disagree on periodicity: two list builds
            if ((rc = eval_pass(ctx, pos_nm, true, false, p2, flags & ~MBPOL_KEEP_LISTS, s, started))) return rc;
            if ((rc = eval_pass(ctx, pos_nm, false, true, p3, flags & ~MBPOL_KEEP_LISTS, s, started))) return rc;
            ctx->lists_valid = false;
        } else {
            if ((rc = eval_pass(ctx, pos_nm, want2, want3, want2 ? p2 : p3, flags, s, started))) return rc;
        }
        if (force) launch_force_scatter(ctx->f_mol.as<double>(), ctx->idx.as<int32_t>(), ctx->n_mol, force, s);
        ctx->f_mol_valid = started;
    }
    if (ctx->timing) cudaEventRecord(ctx->ev[4], s);
    ctx->timed2 = want2 && ctx->stats.n_pairs_owned > 0;
    ctx->timed3 = want3 && ctx->stats.n_triplets_owned > 0;
    if (energy) {
        CKC(cudaMemcpyAsync(energy, ctx->d_energy.ptr, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CKC(cudaStreamSynchronize(s));
    }
    ctx->stats.n_launches = g_launches;
    ctx->stats.lists_sliced = ctx->lists.sliced;
    ctx->stats.n_repartitions = ctx->n_repartitions;
    ctx->stats.n_row_builds = ctx->lists.n_row_builds;
    ctx->stats.lists_reused = ctx->lists.rows_reused;
    if (ctx->timing && !ctx->timing_defer_sync && (rc = mbpol_collect_timing(ctx))) return rc;
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

const double* mbpol_energy_device(const mbpol_ctx* ctx) { return ctx ? ctx->d_energy.as<double>() : nullptr; }

int mbpol_get_energy(mbpol_ctx* ctx, double energy[3], void* stream) {
    if (!ctx || !energy) return fail(ctx, MBPOL_EINVAL, "mbpol_get_energy: null argument");
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaMemcpyAsync(energy, ctx->d_energy.ptr, 3 * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CKC(cudaStreamSynchronize((cudaStream_t)stream));
    return MBPOL_OK;
}

int mbpol_eval_host(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, uint32_t flags) {
    if (!ctx || !pos_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_host: null argument");
    CKC(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->n_particles;
    int rc;
    if ((rc = ctx->pos_stage.reserve(bytes + 8, ctx->err))) return rc;
    if ((rc = ctx->force_stage.reserve(bytes + 8, ctx->err))) return rc;
    cudaStream_t s = 0;
    CKC(cudaMemcpyAsync(ctx->pos_stage.ptr, pos_nm, bytes, cudaMemcpyHostToDevice, s));
    if (force) {
        if (flags & MBPOL_OVERWRITE_FORCES) CKC(cudaMemsetAsync(ctx->force_stage.ptr, 0, bytes, s));
        else CKC(cudaMemcpyAsync(ctx->force_stage.ptr, force, bytes, cudaMemcpyHostToDevice, s));
    }
    if ((rc = mbpol_eval(ctx, ctx->pos_stage.as<double>(), force ? ctx->force_stage.as<double>() : nullptr, energy,
                         flags, (void*)s)))
        return rc;
    if (force) CKC(cudaMemcpyAsync(force, ctx->force_stage.ptr, bytes, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    return MBPOL_OK;
}

int mbpol_eval_onebody(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, int32_t periodic, void* stream) {
    if (!ctx || !pos_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_onebody: null argument");
    if (periodic && !(ctx->box[0] > 0 && ctx->box[1] > 0 && ctx->box[2] > 0))
        return fail(ctx, MBPOL_EBOX, "mbpol_eval_onebody: periodic evaluation needs a box (mbpol_set_box)");
    cudaStream_t s = (cudaStream_t)stream;
    CKC(cudaSetDevice(ctx->device));
    g_launches = 0;
    const int first = (int)((int64_t)ctx->n_mol * ctx->rank / ctx->world);
    const int n = (int)((int64_t)ctx->n_mol * (ctx->rank + 1) / ctx->world) - first;
    int rc;
    if ((rc = ctx->e1.reserve(sizeof(double) * (size_t)(n > 0 ? n : 1), ctx->err))) return rc;
    BoxArg box{{ctx->box[0], ctx->box[1], ctx->box[2]}, periodic ? 1 : 0};
    launch_onebody(pos_nm, ctx->idx.as<int32_t>(), first, n, box, force, ctx->e1.as<double>(), s);
    launch_energy_sum(ctx->e1.as<double>(), n, ctx->e_partial.as<double>(), ctx->d_energy.as<double>() + 2, s);
    if (energy) {
        CKC(cudaMemcpyAsync(energy, ctx->d_energy.as<double>() + 2, sizeof(double), cudaMemcpyDeviceToHost, s));
        CKC(cudaStreamSynchronize(s));
    }
    ctx->stats.n_launches = g_launches;
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

int mbpol_eval_onebody_host(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, int32_t periodic) {
    if (!ctx || !pos_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_onebody_host: null argument");
    CKC(cudaSetDevice(ctx->device));
    const size_t bytes = sizeof(double) * 3 * (size_t)ctx->n_particles;
    int rc;
    if ((rc = ctx->pos_stage.reserve(bytes + 8, ctx->err))) return rc;
    if ((rc = ctx->force_stage.reserve(bytes + 8, ctx->err))) return rc;
    cudaStream_t s = 0;
    CKC(cudaMemcpyAsync(ctx->pos_stage.ptr, pos_nm, bytes, cudaMemcpyHostToDevice, s));
    if (force) CKC(cudaMemcpyAsync(ctx->force_stage.ptr, force, bytes, cudaMemcpyHostToDevice, s));
    if ((rc = mbpol_eval_onebody(ctx, ctx->pos_stage.as<double>(), force ? ctx->force_stage.as<double>() : nullptr,
                                 energy, periodic, (void*)s)))
        return rc;
    if (force) CKC(cudaMemcpyAsync(force, ctx->force_stage.ptr, bytes, cudaMemcpyDeviceToHost, s));
    CKC(cudaStreamSynchronize(s));
    return MBPOL_OK;
}

int mbpol_add_forces_fixed_point(mbpol_ctx* ctx, long long* omm_force, int32_t padded, void* stream) {
    if (!ctx || !omm_force || padded < ctx->n_particles) return fail(ctx, MBPOL_EINVAL, "mbpol_add_forces_fixed_point: bad argument");
    if (!ctx->f_mol_valid) return fail(ctx, MBPOL_EINVAL, "mbpol_add_forces_fixed_point: no evaluated forces");
    CKC(cudaSetDevice(ctx->device));
    launch_force_scatter_fixed(ctx->f_mol.as<double>(), ctx->idx.as<int32_t>(), ctx->n_mol, omm_force, padded,
                               (cudaStream_t)stream);
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

static int get_list(mbpol_ctx* ctx, const DeviceBuf& buf, int64_t count, int width, int32_t* out, size_t* n) {
    if (!ctx || !n) return fail(ctx, MBPOL_EINVAL, "mbpol_get_*: null argument");
    if (!ctx->lists_valid) return fail(ctx, MBPOL_EINVAL, "mbpol_get_*: no tuple lists (call mbpol_eval first)");
    if (out) {
        if (*n < (size_t)count) return fail(ctx, MBPOL_EINVAL, "mbpol_get_*: output buffer too small");
        CKC(cudaSetDevice(ctx->device));
        if (count > 0) CKC(cudaMemcpy(out, buf.ptr, sizeof(int32_t) * width * (size_t)count, cudaMemcpyDeviceToHost));
    }
    *n = (size_t)count;
    return MBPOL_OK;
}

int mbpol_get_pairs(mbpol_ctx* ctx, int32_t* out, size_t* n) {
    if (!ctx) return fail(nullptr, MBPOL_EINVAL, "mbpol_get_pairs: null context");
    return get_list(ctx, ctx->lists.pairs, ctx->lists.n_pairs, 2, out, n);
}
int mbpol_get_triplets(mbpol_ctx* ctx, int32_t* out, size_t* n) {
    if (!ctx) return fail(nullptr, MBPOL_EINVAL, "mbpol_get_triplets: null context");
    return get_list(ctx, ctx->lists.trips, ctx->lists.n_trips, 3, out, n);
}

int mbpol_get_stats(const mbpol_ctx* ctx, mbpol_stats* out) {
    if (!ctx || !out) return MBPOL_EINVAL;
    *out = ctx->stats;
    return MBPOL_OK;
}

int mbpol_set_timing(mbpol_ctx* ctx, int32_t enabled) {
    if (!ctx) return MBPOL_EINVAL;
    ctx->timing = enabled != 0;
    return MBPOL_OK;
}

int mbpol_pip_eval(int order, const double* x, double* out, int64_t n, int64_t ld, void* stream) {
    std::string err;
    int rc = pip_launch(order, x, out, n, ld, (cudaStream_t)stream, &err);
    if (rc) g_create_error = err;
    return rc;
}

int mbpol_pip_instruction_count(int order) { return pip_instruction_count(order); }

int mbpol_eval_tuples(int order, const double* pos_nm, int32_t n_mol, const int32_t* tuples, int64_t n,
                      const double box_nm[3], int periodic, double* e_each, double* f_each, void* stream) {
    mbpol_ctx* ctx = nullptr;
    if ((order != 2 && order != 3) || !pos_nm || !tuples || !e_each || !f_each || n_mol <= 0)
        return fail(nullptr, MBPOL_EINVAL, "mbpol_eval_tuples: bad argument");
    if (n <= 0) return MBPOL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int nv = order == 2 ? NV2 : NV3;
    const int64_t ld = round_up(n, 32);
    DeviceBuf X, O, I;
    std::string err;
    int rc;
    if ((rc = X.reserve(sizeof(double) * nv * ld, err)) ||
        (rc = O.reserve(sizeof(double) * (nv + 1) * ld, err)) || (rc = I.reserve(sizeof(int32_t) * 3 * (size_t)n_mol, err)))
        return fail(nullptr, rc, err);
    std::vector<int32_t> idx(3 * (size_t)n_mol);
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = (int32_t)i;
    CKC(cudaMemcpyAsync(I.ptr, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice, s));
    BoxArg box;
    for (int d = 0; d < 3; ++d) box.len[d] = box_nm ? box_nm[d] : 0.0;
    box.periodic = periodic ? 1 : 0;
    if (order == 2) {
        launch_dimer_vars(pos_nm, I.as<int32_t>(), tuples, 0, n, ld, box, X.as<double>(), s);
        if ((rc = pip_launch(2, X.as<double>(), O.as<double>(), n, ld, s, &err, true))) return fail(nullptr, rc, err);
        launch_dimer_forces(pos_nm, I.as<int32_t>(), tuples, 0, n, ld, box, O.as<double>(), e_each, f_each, s);
    } else {
        launch_trimer_vars(pos_nm, I.as<int32_t>(), tuples, 0, n, ld, box, X.as<double>(), s);
        if ((rc = pip_launch(3, X.as<double>(), O.as<double>(), n, ld, s, &err, true))) return fail(nullptr, rc, err);
        launch_trimer_forces(pos_nm, I.as<int32_t>(), tuples, 0, n, ld, box, O.as<double>(), e_each, f_each, s);
    }
    CKC(cudaStreamSynchronize(s));   // temporaries are freed on return
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

// ---- FP64 FMA-pipe peak ------------------------------------------------------------------------------------------
// 16 independent accumulators per thread, g_i = fma(c_i, a, g_i): two fresh 64-bit register reads + one repeated operand,
// the cheapest operand pattern for the register file (profiles/r01_dfma_operand_microbench.md: 0.91-0.99 of nominal;
// three fresh operands issue only every 3rd cycle).  Inline PTX keeps the compiler from restructuring the loop.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double g[16], c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        g[i] = threadIdx.x + i;
        c[i] = b + 1e-9 * (threadIdx.x & 3) + 1e-8 * i;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(g[i]) : "d"(c[i]), "d"(a));
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += g[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int mbpol_measure_fp64_peak(double* tflops, double* ms_out) {
    mbpol_ctx* ctx = nullptr;
    if (!tflops) return MBPOL_EINVAL;
    int dev = 0, sms = 0;
    CKC(cudaGetDevice(&dev));
    CKC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 4096;   // 16 warps per scheduler
    double* out = nullptr;
    CKC(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1e-6, 0.999999);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    CKC(cudaGetLastError());
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    *tflops = flop / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return MBPOL_OK;
}

}  // extern "C"
