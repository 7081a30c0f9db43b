// OpenMM CUDA-platform buffers behind the C ABI (mbpol_eval_openmm, include/mbpol_b200.h).
//
// The reference's CUDA kernels work directly on OpenMM's own device arrays (platforms/cuda/src/CudaMBPolKernels.cpp:
// 278-304 for 2B, 1976-2005 for 3B; kernels/twobodyForce.cu:442-457 for the force format):
//   cu.getPosq()          real4 per atom (double4 with CudaPrecision=double), in OpenMM's REORDERED atom order
//   cu.getAtomIndex()     slot -> original particle index (OpenMM sorts atoms spatially; the CudaForceInfo of the plugin,
//                         CudaMBPolKernels.cpp:147-162, keeps the three atoms of a molecule one group)
//   cu.getForce()         int64 fixed point, value * 2^32, SoA [x: padded][y: padded][z: padded], slot order
//   cu.getEnergyBuffer()  per-thread partial sums; we add into element 0
// The MBPol*Force objects name ORIGINAL particle indices, so: gather posq into an original-order double3 array (and
// remember where each original particle sits), run the ordinary pipeline, scatter the per-molecule forces to the SLOTS.
#include "context.h"

namespace {

template <class Real4>
__global__ void gather_posq_kernel(const Real4* __restrict__ posq, const int32_t* __restrict__ atom_index, int num_atoms,
                                   int n_particles, double* __restrict__ pos, int32_t* __restrict__ slot_of) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // slot in OpenMM's current order
    if (i >= num_atoms) return;
    const int p = atom_index ? atom_index[i] : i;               // original particle index
    if (p < 0 || p >= n_particles) return;
    const Real4 q = posq[i];
    pos[3 * (size_t)p] = (double)q.x;
    pos[3 * (size_t)p + 1] = (double)q.y;
    pos[3 * (size_t)p + 2] = (double)q.z;
    slot_of[p] = i;
}

// buf[c * padded + slot_of[p]] += (long long)(f * 2^32)   -- one thread per (molecule, component); every particle belongs
// to one molecule (checked at mbpol_create), other OpenMM kernels may add to the same entries concurrently on other
// streams only through the same atomics, so use them too
__global__ void scatter_fixed_slots_kernel(const double* __restrict__ f_mol, const int32_t* __restrict__ idx, int n_mol,
                                           const int32_t* __restrict__ slot_of, long long* __restrict__ buf, int padded) {
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gid >= 9LL * n_mol) return;
    const int m = (int)(gid / 9), c = (int)(gid - 9LL * m);
    const int slot = slot_of[idx[3 * m + c / 3]];
    atomicAdd(reinterpret_cast<unsigned long long*>(buf + (size_t)(c % 3) * padded + slot),
              static_cast<unsigned long long>((long long)(f_mol[gid] * 4294967296.0)));
}

template <class Real>
__global__ void add_energy_kernel(const double* __restrict__ e, int first, int count, Real* __restrict__ energy_buffer) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < count; ++k) tot += e[first + k];
        energy_buffer[0] += (Real)tot;
    }
}

}  // namespace

extern "C" int mbpol_eval_openmm(mbpol_ctx* ctx, const void* posq, int32_t posq_is_double, const int32_t* atom_index,
                                 int32_t num_atoms, long long* force_buffer, int32_t padded_num_atoms,
                                 void* energy_buffer, int32_t energy_is_double, uint32_t flags, void* stream) {
    if (!ctx || !posq) return fail(ctx, MBPOL_EINVAL, "mbpol_eval_openmm: null argument");
    if (num_atoms < ctx->n_particles || (force_buffer && padded_num_atoms < num_atoms))
        return fail(ctx, MBPOL_EINVAL, "mbpol_eval_openmm: atom counts do not match the context");
    cudaStream_t s = (cudaStream_t)stream;
    CKC(cudaSetDevice(ctx->device));
    int rc;
    const size_t np = (size_t)ctx->n_particles;
    if ((rc = ctx->pos_stage.reserve(sizeof(double) * 3 * np + 8, ctx->err))) return rc;
    if ((rc = ctx->omm_slot.reserve(sizeof(int32_t) * np + 4, ctx->err))) return rc;
    const unsigned grid = (unsigned)((num_atoms + 255) / 256);
    {
        NvtxRange r("mbpol:gather_posq");
        if (posq_is_double)
            gather_posq_kernel<double4><<<grid, 256, 0, s>>>(static_cast<const double4*>(posq), atom_index, num_atoms,
                                                             ctx->n_particles, ctx->pos_stage.as<double>(),
                                                             ctx->omm_slot.as<int32_t>());
        else
            gather_posq_kernel<float4><<<grid, 256, 0, s>>>(static_cast<const float4*>(posq), atom_index, num_atoms,
                                                            ctx->n_particles, ctx->pos_stage.as<double>(),
                                                            ctx->omm_slot.as<int32_t>());
    }
    if ((rc = mbpol_eval(ctx, ctx->pos_stage.as<double>(), nullptr, nullptr, flags, s))) return rc;
    ctx->stats.n_launches += 1;
    if (force_buffer && ctx->n_mol > 0 && ctx->f_mol_valid) {
        scatter_fixed_slots_kernel<<<(unsigned)((9LL * ctx->n_mol + 287) / 288), 288, 0, s>>>(
            ctx->f_mol.as<double>(), ctx->idx.as<int32_t>(), ctx->n_mol, ctx->omm_slot.as<int32_t>(), force_buffer,
            padded_num_atoms);
        ctx->stats.n_launches += 1;
    }
    if (energy_buffer) {
        // d_energy = {E_2B, E_3B}; a term that was not requested is zero (mbpol_eval clears both)
        if (energy_is_double) add_energy_kernel<double><<<1, 32, 0, s>>>(ctx->d_energy.as<double>(), 0, 2, static_cast<double*>(energy_buffer));
        else add_energy_kernel<float><<<1, 32, 0, s>>>(ctx->d_energy.as<double>(), 0, 2, static_cast<float*>(energy_buffer));
        ctx->stats.n_launches += 1;
    }
    CKC(cudaGetLastError());
    return MBPOL_OK;
}
