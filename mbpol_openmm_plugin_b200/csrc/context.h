// The context object behind the C ABI and the helpers shared by the translation units that implement it
// (capi.cu: single-GPU pipeline, dist.cu: NCCL collectives, nve.cu: on-device integrator).  Not part of the public ABI.
#pragma once

#include <cstring>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "lists.h"

using namespace mbpol;

struct mbpol_ctx {
    int device = 0;
    int n_mol = 0, n_particles = 0;
    double cut2 = 1e10, cut3 = 1e10;
    int method2 = MBPOL_CUTOFF_NONPERIODIC, method3 = MBPOL_CUTOFF_NONPERIODIC;
    double box[3] = {0, 0, 0};
    int rank = 0, world = 1;
    // multi-GPU: owner range of this rank from the last FULL list build.  While it is valid, list builds are SLICED:
    // only this rank's owners and their halo get neighbour rows (TupleLists::build).  Every rank refreshes it at the
    // same evaluation count, from the same positions, so the ranges always tile [0, n_mol).
    bool slice_valid = false, slice_want2 = false, slice_want3 = false, slice_periodic = false;
    int slice_lo = 0, slice_hi = 0, slice_age = 0, repartition_every = 50;
    int64_t n_repartitions = 0;
    bool lists_valid = false;
    bool timing = false;
    std::string err;
    mbpol_stats stats{};

    DeviceBuf idx;
    TupleLists lists;
    ContributionIndex ci2, ci3;
    DeviceBuf x2, o2, e2, f2, x3, o3, e3, f3;
    DeviceBuf f_mol, e_partial, d_energy, pos_stage, force_stage, e1;
    cudaEvent_t ev[6] = {}, evp[4] = {};      // stage events; events around the two polynomial kernels
    cudaEvent_t evc[4] = {};                  // mbpol_eval_distributed: around the broadcast and the reduce
    bool timing_defer_sync = false;           // the caller reads the stage events itself (mbpol_collect_timing)
    bool timed2 = false, timed3 = false;
    bool f_mol_valid = false;
    // Two body orders on two streams (capi.cu, eval_pass): the 2B chain runs on `side` next to the 3B chain on the caller's
    // stream, so that the partial last wave of one polynomial kernel and the small kernels around it overlap the other
    // chain.  Only when the stage events are off (they describe a serial schedule); MBPOL_SERIAL_ORDERS=1 disables it.
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DeviceBuf cub_tmp2, e_partial2;           // scratch of the chain on `side`
    bool serial_orders = false;
    // multi-GPU (dist.cu)
    void* comm = nullptr;                     // ncclComm_t
    bool comm_owned = false;
    DeviceBuf dist_buf;                       // partial forces of this rank [3 n_particles] + reduced energies [2]
    cudaStream_t side2 = nullptr, copy_stream = nullptr;      // mbpol_eval_distributed_host: compute / copy-back streams
    cudaEvent_t ev_slice[4] = {};
    // OpenMM buffers (omm.cu): original particle -> reordered slot, built by the posq gather
    DeviceBuf omm_slot;
    // Verlet skin (lists.cu) and on-device integrator (nve.cu)
    double skin = 0.0;
    DeviceBuf nve_force, nve_ke, nve_log, nve_msite;
    // dispersion (dispersion.cu): its own molecule-pair list (different cutoff), force blocks and contribution index
    TupleLists dlists;
    ContributionIndex ci_d;
    DeviceBuf e_d, f_d, f_mol_d;
};

// reads the stage events of the last evaluation into ctx->stats (synchronises on the last event)
int mbpol_collect_timing(mbpol_ctx* ctx);

// NVTX ranges around the stages; free when no profiler is attached
struct NvtxRange {
    bool open = true;
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    void end() {
        if (open) nvtxRangePop();
        open = false;
    }
    ~NvtxRange() { end(); }
};

extern thread_local std::string g_create_error;

inline int fail(mbpol_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CKC(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(ctx, MBPOL_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)
