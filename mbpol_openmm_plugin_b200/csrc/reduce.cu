// Deterministic segmented force reduction and energy sums -- no floating-point atomics.
//
// The reference accumulates serially `forces[idx] += ...` in list order (MBPolReferenceTwoBodyForce.cpp:276-290,
// MBPolReferenceThreeBodyForce.cpp:279-287); the CUDA tree it ships uses int64 fixed-point atomics
// (platforms/cuda/src/kernels/twobodyForce.cu:442-457).  Here every tuple writes its own 6x3 / 9x3 force block and
// each molecule then sums the blocks that name it, walking its contribution segment (lists.cu: stable sort by
// molecule of (tuple, role) entries) in ascending (tuple, role) order -> bitwise reproducible run to run.
#include "lists.h"

namespace mbpol {

namespace {

// one thread per (molecule, component c in 0..8 = atom*3 + xyz)
__global__ void force_reduce_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ vals, int width,
                                    const double* __restrict__ f_tuple, int n_mol, double* __restrict__ f_mol,
                                    int accumulate) {
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gid >= 9LL * n_mol) return;
    const int m = (int)(gid / 9), c = (int)(gid - 9LL * m);
    double acc = 0.0;
    for (int64_t k = off[m]; k < off[m + 1]; ++k) {
        const int32_t e = __ldg(vals + k);            // local tuple * width + role
        acc += __ldg(f_tuple + 9 * (size_t)e + c);    // block (tuple, role) starts at 9*(tuple*width + role)
    }
    f_mol[gid] = accumulate ? f_mol[gid] + acc : acc;
}

__global__ void force_scatter_kernel(const double* __restrict__ f_mol, const int32_t* __restrict__ idx, int n_mol,
                                     double* __restrict__ force) {
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gid >= 9LL * n_mol) return;
    const int m = (int)(gid / 9), c = (int)(gid - 9LL * m);
    const int p = idx[3 * m + c / 3];
    force[3 * (size_t)p + c % 3] += f_mol[gid];
}

__global__ void force_scatter_fixed_kernel(const double* __restrict__ f_mol, const int32_t* __restrict__ idx, int n_mol,
                                           long long* __restrict__ buf, int padded) {
    const int64_t gid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (gid >= 9LL * n_mol) return;
    const int m = (int)(gid / 9), c = (int)(gid - 9LL * m);
    const int p = idx[3 * m + c / 3];
    // same conversion as the plugin's CUDA kernels: (long long)(f * 0x100000000)
    atomicAdd(reinterpret_cast<unsigned long long*>(buf + (size_t)(c % 3) * padded + p),
              static_cast<unsigned long long>((long long)(f_mol[gid] * 4294967296.0)));
}

constexpr int ESUM_BLOCK = 256, ESUM_MAX_BLOCKS = 1024;

// fixed-shape block sum: warp-shuffle tree inside every warp, then the (<= 32) warp totals by the first warp; the
// order of additions depends on nothing but the block size -> bitwise reproducible
__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;                                            // lane 0 holds the total
}

__device__ __forceinline__ double block_tree_sum(double v, double* sh /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_tree_sum(v);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double tot = 0.0;
    if (warp == 0) {
        tot = warp_tree_sum(lane < nwarp ? sh[lane] : 0.0);
        if (lane == 0) sh[0] = tot;
    }
    __syncthreads();
    return sh[0];
}

__global__ void energy_partial_kernel(const double* __restrict__ e, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += e[i];
    const double tot = block_tree_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

__global__ void energy_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    __shared__ double sh[32];
    const double tot = block_tree_sum(threadIdx.x < nb ? partial[threadIdx.x] : 0.0, sh);
    if (threadIdx.x == 0) *out = tot;
}

inline unsigned blocks(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

}  // namespace

void launch_force_reduce(const ContributionIndex& ci, int width, const double* f_tuple, int n_mol, double* f_mol,
                         bool accumulate, cudaStream_t s) {
    if (n_mol <= 0) return;
    force_reduce_kernel<<<blocks(9LL * n_mol, 288), 288, 0, s>>>(ci.off.as<int64_t>(), ci.vals.as<int32_t>(), width,
                                                                  f_tuple, n_mol, f_mol, accumulate ? 1 : 0); MBPOL_COUNT_LAUNCH(1);
}

void launch_force_scatter(const double* f_mol, const int32_t* idx, int n_mol, double* force, cudaStream_t s) {
    if (n_mol <= 0) return;
    force_scatter_kernel<<<blocks(9LL * n_mol, 288), 288, 0, s>>>(f_mol, idx, n_mol, force); MBPOL_COUNT_LAUNCH(1);
}

void launch_force_scatter_fixed(const double* f_mol, const int32_t* idx, int n_mol, long long* buf, int padded,
                                cudaStream_t s) {
    if (n_mol <= 0) return;
    force_scatter_fixed_kernel<<<blocks(9LL * n_mol, 288), 288, 0, s>>>(f_mol, idx, n_mol, buf, padded); MBPOL_COUNT_LAUNCH(1);
}

void launch_energy_sum(const double* e, int64_t n, double* partial, double* out, cudaStream_t s) {
    int nb = (int)((n + ESUM_BLOCK - 1) / ESUM_BLOCK);
    if (nb > ESUM_MAX_BLOCKS) nb = ESUM_MAX_BLOCKS;
    if (nb < 1) nb = 1;
    energy_partial_kernel<<<nb, ESUM_BLOCK, 0, s>>>(e, n, partial); MBPOL_COUNT_LAUNCH(1);
    energy_final_kernel<<<1, ESUM_MAX_BLOCKS, 0, s>>>(partial, nb, out); MBPOL_COUNT_LAUNCH(1);
}

}  // namespace mbpol
