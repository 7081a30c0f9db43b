// Multi-GPU behind the C ABI: NCCL broadcast of the positions, owner-partitioned evaluation, NCCL reduce of forces +
// energies (BASELINE.json north_star (3)).  One process (or thread) per GPU; every rank holds one mbpol_ctx.
//
//   mbpol_comm_unique_id / mbpol_comm_init_rank   bootstrap: rank 0 creates the 128-byte NCCL id, the host program ships
//                                                it to the other ranks by whatever means it has (MPI, a file, a socket,
//                                                torch.distributed's store), every rank joins
//   mbpol_comm_adopt                             alternative: use a communicator the host program already owns
//   mbpol_eval_distributed                       broadcast -> lists + 2B + 3B of this rank's owner range -> reduce
//
// The data path between the two collectives has no exchange step (tuples are independent).  The force array and the two
// energies leave in ONE grouped NCCL operation (one launch); the partial forces are scattered straight into the send
// buffer, nothing is staged through the host.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy already loaded into the process if there is one, e.g.
// PyTorch's), so libmbpol_b200.so keeps loading on machines without NCCL; the prototypes below are the stable C ABI of
// nccl.h (NCCL 2.x).
#include <dlfcn.h>

#include <mutex>

#include "context.h"

namespace {

struct NcclUniqueId {
    char internal[MBPOL_COMM_ID_BYTES];
};
typedef void* NcclComm;
enum { NCCL_SUM = 0, NCCL_FLOAT64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*CommCount)(NcclComm, int*) = nullptr;
    int (*CommUserRank)(NcclComm, int*) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
    std::string why;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

const NcclApi& nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.ok || !g_nccl.why.empty()) return g_nccl;
    void* h = nullptr;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        if ((h = dlopen(name, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL))) break;      // the copy the process already has
        if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) {
        g_nccl.why = std::string("NCCL is not available: ") + (dlerror() ? dlerror() : "libnccl.so.2 not found");
        return g_nccl;
    }
    bool all = true;
    auto sym = [&](const char* n, auto& fn) {
        void* p = dlsym(h, n);
        if (!p) all = false;
        fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
    };
    sym("ncclGetUniqueId", g_nccl.GetUniqueId);
    sym("ncclCommInitRank", g_nccl.CommInitRank);
    sym("ncclCommDestroy", g_nccl.CommDestroy);
    sym("ncclCommCount", g_nccl.CommCount);
    sym("ncclCommUserRank", g_nccl.CommUserRank);
    sym("ncclBroadcast", g_nccl.Broadcast);
    sym("ncclReduce", g_nccl.Reduce);
    sym("ncclAllReduce", g_nccl.AllReduce);
    sym("ncclGroupStart", g_nccl.GroupStart);
    sym("ncclGroupEnd", g_nccl.GroupEnd);
    sym("ncclGetErrorString", g_nccl.GetErrorString);
    if (!all) g_nccl.why = "NCCL library lacks a required symbol";
    g_nccl.ok = all;
    return g_nccl;
}

#define CKN(call)                                                                                                    \
    do {                                                                                                             \
        int r_ = (call);                                                                                             \
        if (r_ != 0) return fail(ctx, MBPOL_ECOMM, std::string(#call) + ": " + N.GetErrorString(r_));                \
    } while (0)

}  // namespace

extern "C" {

int mbpol_comm_unique_id(void* id_out) {
    mbpol_ctx* ctx = nullptr;
    if (!id_out) return fail(nullptr, MBPOL_EINVAL, "mbpol_comm_unique_id: null argument");
    const NcclApi& N = nccl();
    if (!N.ok) return fail(nullptr, MBPOL_ECOMM, N.why);
    NcclUniqueId id;
    CKN(N.GetUniqueId(&id));
    memcpy(id_out, id.internal, MBPOL_COMM_ID_BYTES);
    return MBPOL_OK;
}

int mbpol_comm_init_rank(mbpol_ctx* ctx, const void* id_bytes, int32_t rank, int32_t world) {
    if (!ctx || !id_bytes || world < 1 || rank < 0 || rank >= world)
        return fail(ctx, MBPOL_EINVAL, "mbpol_comm_init_rank: bad argument");
    const NcclApi& N = nccl();
    if (!N.ok) return fail(ctx, MBPOL_ECOMM, N.why);
    int rc;
    if ((rc = mbpol_comm_destroy(ctx))) return rc;
    CKC(cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(id.internal, id_bytes, MBPOL_COMM_ID_BYTES);
    NcclComm comm = nullptr;
    // This communicator carries two messages per step: 3 n_particles doubles out (broadcast) and back (reduce), 19 MB at
    // 262,144 waters.  NCCL's cost model picks the Simple protocol for them; measured on 8 B200 over NVSwitch the reduce
    // then takes 0.25 ms, with LL128 0.06 ms (profiles/r02_nccl_proto_8gpu.txt; between two ranks Simple is the faster
    // one, 0.12 against 0.16 ms, so the default is only changed from four ranks on).  NCCL reads NCCL_PROTO when a
    // communicator is tuned, so it is set for the duration of THIS init only -- unless the user set it, or asks for
    // something else through MBPOL_NCCL_PROTO (empty string: leave NCCL alone).
    const char* want = getenv("MBPOL_NCCL_PROTO");
    if (!want) want = "LL128";
    const bool set_proto = !getenv("NCCL_PROTO") && want[0] != 0 && (world >= 4 || getenv("MBPOL_NCCL_PROTO"));
    if (set_proto) setenv("NCCL_PROTO", want, 1);
    const int init_rc = N.CommInitRank(&comm, world, id, rank);
    if (set_proto) unsetenv("NCCL_PROTO");
    CKN(init_rc);
    ctx->comm = comm;
    ctx->comm_owned = true;
    return mbpol_set_partition(ctx, rank, world);
}

int mbpol_comm_adopt(mbpol_ctx* ctx, void* nccl_comm) {
    if (!ctx || !nccl_comm) return fail(ctx, MBPOL_EINVAL, "mbpol_comm_adopt: null argument");
    const NcclApi& N = nccl();
    if (!N.ok) return fail(ctx, MBPOL_ECOMM, N.why);
    int rc, rank = 0, world = 1;
    if ((rc = mbpol_comm_destroy(ctx))) return rc;
    CKN(N.CommCount(nccl_comm, &world));
    CKN(N.CommUserRank(nccl_comm, &rank));
    ctx->comm = nccl_comm;
    ctx->comm_owned = false;
    return mbpol_set_partition(ctx, rank, world);
}

int mbpol_comm_destroy(mbpol_ctx* ctx) {
    if (!ctx) return MBPOL_EINVAL;
    if (ctx->comm && ctx->comm_owned) {
        const NcclApi& N = nccl();
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        if (N.ok) N.CommDestroy(ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->comm_owned = false;
    return MBPOL_OK;
}

int mbpol_eval_distributed(mbpol_ctx* ctx, double* pos_nm, double* force, double* energy, uint32_t flags, void* stream) {
    if (!ctx || !pos_nm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_distributed: null argument");
    if (!ctx->comm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_distributed: no communicator (mbpol_comm_init_rank)");
    const NcclApi& N = nccl();
    if (!N.ok) return fail(ctx, MBPOL_ECOMM, N.why);
    cudaStream_t s = (cudaStream_t)stream;
    CKC(cudaSetDevice(ctx->device));
    const size_t n3 = 3 * (size_t)ctx->n_particles;
    const bool all = flags & MBPOL_DIST_ALLREDUCE, receiver = all || ctx->rank == 0;
    int rc;
    if ((rc = ctx->dist_buf.reserve(sizeof(double) * (n3 + 8), ctx->err))) return rc;
    double* part = ctx->dist_buf.as<double>();              // this rank's partial forces [n3], reduced energies [2]
    if (receiver && !force) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_distributed: the receiving rank needs a force buffer");
    if (ctx->timing) cudaEventRecord(ctx->evc[0], s);
    if (!(flags & MBPOL_DIST_NO_BCAST)) {
        NvtxRange r("mbpol:broadcast");
        CKN(N.Broadcast(pos_nm, pos_nm, n3, NCCL_FLOAT64, 0, ctx->comm, s));
    }
    if (ctx->timing) cudaEventRecord(ctx->evc[1], s);
    CKC(cudaMemsetAsync(part, 0, sizeof(double) * n3, s));
    const bool timing = ctx->timing;
    ctx->timing_defer_sync = true;                          // stage events are read once, after the reduce
    rc = mbpol_eval(ctx, pos_nm, part, nullptr, flags & (MBPOL_DO_2B | MBPOL_DO_3B | MBPOL_KEEP_LISTS), s);
    ctx->timing_defer_sync = false;
    if (rc) return rc;
    if (timing) cudaEventRecord(ctx->evc[2], s);
    NvtxRange r_red("mbpol:reduce");
    double* e_sum = part + n3;
    CKN(N.GroupStart());                                    // forces and energies in one NCCL launch
    if (all) {
        CKN(N.AllReduce(part, force, n3, NCCL_FLOAT64, NCCL_SUM, ctx->comm, s));
        CKN(N.AllReduce(ctx->d_energy.ptr, e_sum, 2, NCCL_FLOAT64, NCCL_SUM, ctx->comm, s));
    } else {
        CKN(N.Reduce(part, force, n3, NCCL_FLOAT64, NCCL_SUM, 0, ctx->comm, s));
        CKN(N.Reduce(ctx->d_energy.ptr, e_sum, 2, NCCL_FLOAT64, NCCL_SUM, 0, ctx->comm, s));
    }
    CKN(N.GroupEnd());
    r_red.end();
    if (timing) cudaEventRecord(ctx->evc[3], s);
    if (energy && receiver) {
        CKC(cudaMemcpyAsync(energy, e_sum, 2 * sizeof(double), cudaMemcpyDeviceToHost, s));
        CKC(cudaStreamSynchronize(s));
    }
    if (timing) {
        CKC(cudaEventSynchronize(ctx->evc[3]));
        if ((rc = mbpol_collect_timing(ctx))) return rc;
        cudaEventElapsedTime(&ctx->stats.ms_bcast, ctx->evc[0], ctx->evc[1]);
        cudaEventElapsedTime(&ctx->stats.ms_comm_reduce, ctx->evc[2], ctx->evc[3]);
    }
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

int mbpol_eval_distributed_host(mbpol_ctx* ctx, const double* pos_host, double* force_host, double* energy, uint32_t flags) {
    if (!ctx) return MBPOL_EINVAL;
    if (!ctx->comm) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_distributed_host: no communicator (mbpol_comm_init_rank)");
    if (ctx->rank == 0 && (!pos_host || !force_host))
        return fail(ctx, MBPOL_EINVAL, "mbpol_eval_distributed_host: rank 0 needs the host position and force buffers");
    const NcclApi& N = nccl();
    if (!N.ok) return fail(ctx, MBPOL_ECOMM, N.why);
    CKC(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->side2 ? ctx->side2 : nullptr;        // compute / collectives
    if (!s) {
        CKC(cudaStreamCreateWithFlags(&ctx->side2, cudaStreamNonBlocking));
        CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (auto& e : ctx->ev_slice) CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s = ctx->side2;
    }
    const size_t n3 = 3 * (size_t)ctx->n_particles, bytes = sizeof(double) * n3;
    int rc;
    if ((rc = ctx->pos_stage.reserve(bytes + 8, ctx->err))) return rc;
    if ((rc = ctx->force_stage.reserve(bytes + 8, ctx->err))) return rc;
    if ((rc = ctx->dist_buf.reserve(sizeof(double) * (n3 + 8), ctx->err))) return rc;
    double* pos = ctx->pos_stage.as<double>();
    double* part = ctx->dist_buf.as<double>();
    double* sum = ctx->force_stage.as<double>();
    if (ctx->rank == 0) CKC(cudaMemcpyAsync(pos, pos_host, bytes, cudaMemcpyHostToDevice, s));
    {
        NvtxRange r("mbpol:broadcast");
        CKN(N.Broadcast(pos, pos, n3, NCCL_FLOAT64, 0, ctx->comm, s));
    }
    CKC(cudaMemsetAsync(part, 0, bytes, s));
    const bool timing = ctx->timing;
    ctx->timing = false;                                    // the stage events describe mbpol_eval_distributed, not this path
    rc = mbpol_eval(ctx, pos, part, nullptr, flags & (MBPOL_DO_2B | MBPOL_DO_3B | MBPOL_KEEP_LISTS), s);
    ctx->timing = timing;
    if (rc) return rc;
    NvtxRange r_red("mbpol:reduce+d2h");
    double* e_sum = part + n3;
    constexpr int NSLICE = 4;
    const size_t per = ((n3 + NSLICE - 1) / NSLICE + 1) & ~(size_t)1;
    for (int c = 0; c < NSLICE; ++c) {
        const size_t lo = (size_t)c * per, hi = lo + per < n3 ? lo + per : n3;
        if (lo >= hi) break;
        CKN(N.GroupStart());
        CKN(N.Reduce(part + lo, sum + lo, hi - lo, NCCL_FLOAT64, NCCL_SUM, 0, ctx->comm, s));
        if (c == 0) CKN(N.Reduce(ctx->d_energy.ptr, e_sum, 2, NCCL_FLOAT64, NCCL_SUM, 0, ctx->comm, s));
        CKN(N.GroupEnd());
        if (ctx->rank == 0) {
            CKC(cudaEventRecord(ctx->ev_slice[c], s));
            CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_slice[c], 0));
            CKC(cudaMemcpyAsync(force_host + lo, sum + lo, sizeof(double) * (hi - lo), cudaMemcpyDeviceToHost, ctx->copy_stream));
            if (c == 0 && energy)
                CKC(cudaMemcpyAsync(energy, e_sum, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
    }
    r_red.end();
    CKC(cudaStreamSynchronize(s));
    if (ctx->rank == 0) CKC(cudaStreamSynchronize(ctx->copy_stream));
    CKC(cudaGetLastError());
    return MBPOL_OK;
}

}  // extern "C"
