// GPU cell list and tuple enumeration: oxygens -> sorted neighbour rows -> compacted, canonically ordered
// pair list (I > J) and chain-triplet list (I > J > K), plus the tuple->molecule contribution index used by the
// deterministic force reduction.
//
// Set semantics reproduced bit-exactly:
//   pairs     {(I,J) : I > J, d2(I,J) <= c2^2}                       OpenMM computeNeighborListVoxelHash as called at
//                                                                    platforms/reference/src/MBPolReferenceKernels.cpp:380-382
//   triplets  {(I,J,K) : J in nearby[I], K in nearby[J]},  nearby[i] = {j < i : d2(i,j) <= c3^2}
//                                                                    platforms/reference/src/ReferenceThreeNeighborList.cpp:181-226
//   d2 uses periodicDifference / compPairDistanceSquared verbatim in rounding behaviour (:14-33): every operation is
//   issued through a round-to-nearest intrinsic so no FMA contraction can flip a borderline `<=` decision.
// Emission order is canonical (ascending I, then J, then K); the reference's own order is voxel-scan dependent and is
// compared after sorting.
#include <cub/cub.cuh>

#include "lists.h"

namespace mbpol {

namespace {

__device__ __forceinline__ double periodic_difference(double v1, double v2, double period) {
    const double diff = __dsub_rn(v1, v2);
    const double base = __dmul_rn(floor(__dadd_rn(__ddiv_rn(diff, period), 0.5)), period);
    return __dsub_rn(diff, base);
}

// squared distance "from pos1 to pos2" exactly as the reference evaluates it with pos1 = higher index
__device__ __forceinline__ double pair_d2(const double* __restrict__ p1, const double* __restrict__ p2, const GridArg& g) {
    double dx, dy, dz;
    if (!g.periodic) {
        dx = __dsub_rn(p2[0], p1[0]);
        dy = __dsub_rn(p2[1], p1[1]);
        dz = __dsub_rn(p2[2], p1[2]);
    } else {
        dx = periodic_difference(p2[0], p1[0], g.box[0]);
        dy = periodic_difference(p2[1], p1[1], g.box[1]);
        dz = periodic_difference(p2[2], p1[2], g.box[2]);
    }
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Cheap screening distance (reciprocal multiply instead of the three IEEE divisions, FMA contraction allowed): differs
// from pair_d2 by a few ulp at most.  Decisions `d2 <= c^2` are taken from it unless it lands within 1e-9 (relative)
// of a cutoff, in which case the exact reference arithmetic decides -- so the lists stay bit-exact while the divisions
// (3 x ~25 FP64 instructions per candidate, ~90 % of the list-build time) disappear from the common path.
__device__ __forceinline__ double pair_d2_fast(const double* __restrict__ p1, const double* __restrict__ p2,
                                               const GridArg& g) {
    double dx = p2[0] - p1[0], dy = p2[1] - p1[1], dz = p2[2] - p1[2];
    if (g.periodic) {
        dx -= floor(dx * g.inv_box[0] + 0.5) * g.box[0];
        dy -= floor(dy * g.inv_box[1] + 0.5) * g.box[1];
        dz -= floor(dz * g.inv_box[2] + 0.5) * g.box[2];
    }
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ double screened_d2(const double* __restrict__ p1, const double* __restrict__ p2,
                                              const GridArg& g, double c2sq, double c3sq) {
    const double d2 = pair_d2_fast(p1, p2, g);
    const bool near2 = fabs(d2 - c2sq) <= 1e-9 * c2sq, near3 = fabs(d2 - c3sq) <= 1e-9 * c3sq;
    return (near2 || near3) ? pair_d2(p1, p2, g) : d2;
}

__device__ __forceinline__ int cell_coord(double x, int d, const GridArg& g) {
    double f;
    if (g.periodic) {
        f = x / g.box[d];
        f -= floor(f);
    } else {
        f = (x - g.lo[d]) * g.inv_extent[d];
    }
    int c = (int)(f * g.nc[d]);
    return min(max(c, 0), g.nc[d] - 1);
}

__global__ void gather_oxygens_kernel(const double* __restrict__ pos, const int32_t* __restrict__ idx, int n,
                                      double* __restrict__ posO) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const double* p = pos + 3 * (size_t)idx[3 * m];
    posO[3 * m] = p[0];
    posO[3 * m + 1] = p[1];
    posO[3 * m + 2] = p[2];
}

// single-block bounding box (non-periodic systems are small clusters)
__global__ void bbox_kernel(const double* __restrict__ posO, int n, double* __restrict__ out /*lo[3],hi[3]*/) {
    __shared__ double slo[3][256], shi[3][256];
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int m = threadIdx.x; m < n; m += blockDim.x)
        for (int d = 0; d < 3; ++d) {
            lo[d] = fmin(lo[d], posO[3 * m + d]);
            hi[d] = fmax(hi[d], posO[3 * m + d]);
        }
    for (int d = 0; d < 3; ++d) {
        slo[d][threadIdx.x] = lo[d];
        shi[d][threadIdx.x] = hi[d];
    }
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int d = 0; d < 3; ++d) {
                slo[d][threadIdx.x] = fmin(slo[d][threadIdx.x], slo[d][threadIdx.x + s]);
                shi[d][threadIdx.x] = fmax(shi[d][threadIdx.x], shi[d][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0)
        for (int d = 0; d < 3; ++d) {
            out[d] = slo[d][0];
            out[3 + d] = shi[d][0];
        }
}

__global__ void cell_count_kernel(const double* __restrict__ posO, int n, GridArg g, int* __restrict__ cell_of,
                                  int* __restrict__ cell_count) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int cx = cell_coord(posO[3 * m], 0, g), cy = cell_coord(posO[3 * m + 1], 1, g),
              cz = cell_coord(posO[3 * m + 2], 2, g);
    const int c = (cx * g.nc[1] + cy) * g.nc[2] + cz;
    cell_of[m] = c;
    atomicAdd(cell_count + c, 1);      // integer counts: the result does not depend on arrival order
}

__global__ void cell_fill_kernel(int n, const int* __restrict__ cell_of, const int* __restrict__ cell_start,
                                 int* __restrict__ cell_cursor, int* __restrict__ cell_members) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int c = cell_of[m];
    // order inside a cell is arrival order; every consumer sorts its neighbour row by index afterwards
    cell_members[cell_start[c] + atomicAdd(cell_cursor + c, 1)] = m;
}

// Visit all j != i in the 27 (or fewer, for tiny grids) cells around i.
template <class F>
__device__ __forceinline__ void for_each_candidate(int i, const double* __restrict__ posO, const GridArg& g,
                                                   const int* __restrict__ cell_start,
                                                   const int* __restrict__ cell_members, F&& f) {
    int c[3], lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c[d] = cell_coord(posO[3 * i + d], d, g);
        if (g.periodic) {
            if (g.nc[d] >= 3) { lo[d] = c[d] - 1; hi[d] = c[d] + 1; }
            else { lo[d] = 0; hi[d] = g.nc[d] - 1; }          // 1 or 2 cells: visit each exactly once
        } else {
            lo[d] = max(c[d] - 1, 0);
            hi[d] = min(c[d] + 1, g.nc[d] - 1);
        }
    }
    for (int x = lo[0]; x <= hi[0]; ++x) {
        const int wx = (x + g.nc[0]) % g.nc[0];
        for (int y = lo[1]; y <= hi[1]; ++y) {
            const int wy = (y + g.nc[1]) % g.nc[1];
            for (int z = lo[2]; z <= hi[2]; ++z) {
                const int wz = (z + g.nc[2]) % g.nc[2];
                const int cell = (wx * g.nc[1] + wy) * g.nc[2] + wz;
                for (int k = cell_start[cell]; k < cell_start[cell + 1]; ++k) {
                    const int j = cell_members[k];
                    if (j != i) f(j);
                }
            }
        }
    }
}

// pass 1: number of neighbours within c2 / c3 (full rows: both j < i and j > i)
__global__ void row_count_kernel(const double* __restrict__ posO, int n, GridArg g, const int* __restrict__ cell_start,
                                 const int* __restrict__ cell_members, double c2sq, double c3sq,
                                 int* __restrict__ cnt2, int* __restrict__ cnt3) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = cell_members[t];      // cell order: the lanes of a warp scan the same candidate cells together
    int n2 = 0, n3 = 0;
    for_each_candidate(i, posO, g, cell_start, cell_members, [&](int j) {
        const int hi = max(i, j), lo = min(i, j);
        const double d2 = screened_d2(posO + 3 * hi, posO + 3 * lo, g, c2sq, c3sq);
        n2 += (d2 <= c2sq);
        n3 += (d2 <= c3sq);
    });
    cnt2[i] = n2;
    cnt3[i] = n3;
}

__device__ __forceinline__ int sort_row(int* __restrict__ row, int len, int self) {
    for (int a = 1; a < len; ++a) {          // rows are short (tens of entries): insertion sort in place
        const int v = row[a];
        int b = a - 1;
        while (b >= 0 && row[b] > v) {
            row[b + 1] = row[b];
            --b;
        }
        row[b + 1] = v;
    }
    int lower = 0;
    while (lower < len && row[lower] < self) ++lower;
    return lower;
}

// pass 2: fill + sort the rows, record how many entries are lower-index ("nearby" in the reference's words)
__global__ void row_fill_kernel(const double* __restrict__ posO, int n, GridArg g, const int* __restrict__ cell_start,
                                const int* __restrict__ cell_members, double c2sq, double c3sq,
                                const int64_t* __restrict__ ro2, const int64_t* __restrict__ ro3,
                                int* __restrict__ nbr2, int* __restrict__ nbr3, int* __restrict__ low2,
                                int* __restrict__ low3) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int i = cell_members[t];
    int* r2 = nbr2 + ro2[i];
    int* r3 = nbr3 + ro3[i];
    int n2 = 0, n3 = 0;
    for_each_candidate(i, posO, g, cell_start, cell_members, [&](int j) {
        const int hi = max(i, j), lo = min(i, j);
        const double d2 = screened_d2(posO + 3 * hi, posO + 3 * lo, g, c2sq, c3sq);
        if (d2 <= c2sq) r2[n2++] = j;
        if (d2 <= c3sq) r3[n3++] = j;
    });
    low2[i] = sort_row(r2, n2, i);
    low3[i] = sort_row(r3, n3, i);
}

// Fast path: one warp per molecule, ONE pass.  Only lower-index neighbours are kept (nothing downstream reads the upper
// part of a row); candidates are compacted with ballot/popc into a shared-memory staging row, rank-sorted by the warp and
// written to fixed-stride rows (ROW_CAP2 / ROW_CAP3 entries).  A row that does not fit raises `overflow` and the caller
// falls back to the two-pass CSR path above -- same lists either way.
constexpr int ROW_CAP2 = 64, ROW_CAP3 = 32, ROWS_WARPS = 8;
constexpr int SKIN_CAP = 128;          // super rows (cutoff + skin): up to ~59 lower-index neighbours at 0.75 nm in liquid water


template <int CAP2>
__global__ void __launch_bounds__(ROWS_WARPS * 32) rows_warp_kernel(
    const double* __restrict__ posO, int n, GridArg g, const int* __restrict__ cell_start,
    const int* __restrict__ cell_members, double c2sq, double c3sq, int* __restrict__ nbr2, int* __restrict__ nbr3,
    int* __restrict__ low2, int* __restrict__ low3, int64_t* __restrict__ ro2, int64_t* __restrict__ ro3,
    int* __restrict__ overflow, int i_lo, int i_hi, const unsigned char* __restrict__ halo) {
    __shared__ int s2[ROWS_WARPS][CAP2], s3[ROWS_WARPS][ROW_CAP3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * ROWS_WARPS + warp;
    if (w >= n) return;
    const int i = cell_members[w];
    // sliced build: rows of the owner range [i_lo, i_hi), or (second launch) of the flagged halo molecules below it
    if (halo ? !(i < i_lo && halo[i]) : (i < i_lo || i >= i_hi)) return;
    const unsigned lt = (1u << lane) - 1u;
    int c[3], lo[3], hi[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        c[d] = cell_coord(posO[3 * i + d], d, g);
        if (g.periodic) {
            if (g.nc[d] >= 3) { lo[d] = c[d] - 1; hi[d] = c[d] + 1; }
            else { lo[d] = 0; hi[d] = g.nc[d] - 1; }
        } else {
            lo[d] = max(c[d] - 1, 0);
            hi[d] = min(c[d] + 1, g.nc[d] - 1);
        }
    }
    int n2 = 0, n3 = 0;
    auto scan = [&](int beg, int end) {                 // members [beg, end) of cell_members, 32 candidates at a time
        for (int k0 = beg; k0 < end; k0 += 32) {
            const int k = k0 + lane;
            const int j = k < end ? cell_members[k] : -1;
            bool p2 = false, p3 = false;
            if (j >= 0 && j < i) {                      // distance from the higher index, like the reference
                const double d2 = screened_d2(posO + 3 * i, posO + 3 * j, g, c2sq, c3sq);
                p2 = d2 <= c2sq;
                p3 = d2 <= c3sq;
            }
            const unsigned b2 = __ballot_sync(0xffffffffu, p2), b3 = __ballot_sync(0xffffffffu, p3);
            if (p2) {
                const int pos = n2 + __popc(b2 & lt);
                if (pos < CAP2) s2[warp][pos] = j;
            }
            if (p3) {
                const int pos = n3 + __popc(b3 & lt);
                if (pos < ROW_CAP3) s3[warp][pos] = j;
            }
            n2 += __popc(b2);
            n3 += __popc(b3);
        }
    };
    for (int x = lo[0]; x <= hi[0]; ++x) {
        const int wx = (x + g.nc[0]) % g.nc[0];
        for (int y = lo[1]; y <= hi[1]; ++y) {
            const int wy = (y + g.nc[1]) % g.nc[1];
            const int base = (wx * g.nc[1] + wy) * g.nc[2];
            if (lo[2] >= 0 && hi[2] < g.nc[2]) {
                // z is the fastest cell index: the (up to three) z-neighbours are ONE contiguous member range, so the
                // warp scans ~28 candidates per step instead of ~9 (the rows are rank-sorted below: order is free)
                scan(cell_start[base + lo[2]], cell_start[base + hi[2] + 1]);
            } else {
                for (int z = lo[2]; z <= hi[2]; ++z) {  // the run wraps around the periodic boundary
                    const int wz = (z + g.nc[2]) % g.nc[2];
                    scan(cell_start[base + wz], cell_start[base + wz + 1]);
                }
            }
        }
    }
    if (n2 > CAP2 || n3 > ROW_CAP3) {
        if (lane == 0) atomicExch(overflow, 1);
        return;
    }
    __syncwarp();
    for (int e = lane; e < n2; e += 32) {               // rank sort: all entries are distinct
        const int v = s2[warp][e];
        int r = 0;
        for (int t = 0; t < n2; ++t) r += s2[warp][t] < v;
        nbr2[(size_t)i * CAP2 + r] = v;
    }
    for (int e = lane; e < n3; e += 32) {
        const int v = s3[warp][e];
        int r = 0;
        for (int t = 0; t < n3; ++t) r += s3[warp][t] < v;
        nbr3[(size_t)i * ROW_CAP3 + r] = v;
    }
    if (lane == 0) {
        low2[i] = n2;
        low3[i] = n3;
        ro2[i] = (int64_t)i * CAP2;
        ro3[i] = (int64_t)i * ROW_CAP3;
    }
}

// ---- Verlet skin ---------------------------------------------------------------------------------------------------
// The reference rebuilds its neighbour lists on every call (MBPolReferenceKernels.cpp:379-383).  With a skin s > 0 the
// cell scan instead produces SUPER rows -- lower-index oxygens within (largest cutoff + s) -- which stay valid while no
// oxygen has moved farther than s / 2 from where it was when they were built; every evaluation filters them with the
// exact d2 <= c^2 test at the CURRENT positions, so the pair / triplet lists the polynomial stage consumes are still the
// reference's sets, bit for bit, at every step.

// flag = 1 if any oxygen moved more than sqrt(limit2) (minimum image) away from its reference position
__global__ void displacement_kernel(const double* __restrict__ posO, const double* __restrict__ ref, int n, GridArg g,
                                    double limit2, int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d2 = pair_d2(ref + 3 * i, posO + 3 * i, g);
    if (!(d2 <= limit2)) atomicExch(flag, 1);
}

// one warp per molecule: keep the entries of its super row that are within the cutoffs NOW (same screened distance test
// as rows_warp_kernel, evaluated from the higher index); order is preserved, so the rows stay sorted
__global__ void __launch_bounds__(ROWS_WARPS * 32) filter_rows_kernel(
    const double* __restrict__ posO, int n, GridArg g, const int* __restrict__ snbr, const int* __restrict__ slow,
    double c2sq, double c3sq, int* __restrict__ nbr2, int* __restrict__ nbr3, int* __restrict__ low2,
    int* __restrict__ low3, int64_t* __restrict__ ro2, int64_t* __restrict__ ro3, int* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * ROWS_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    const unsigned lt = (1u << lane) - 1u;
    const int* row = snbr + (size_t)i * SKIN_CAP;
    const int len = slow[i];
    int n2 = 0, n3 = 0;
    for (int k0 = 0; k0 < len; k0 += 32) {
        const int k = k0 + lane;
        const int j = k < len ? row[k] : -1;
        bool p2 = false, p3 = false;
        if (j >= 0) {
            const double d2 = screened_d2(posO + 3 * i, posO + 3 * j, g, c2sq, c3sq);
            p2 = d2 <= c2sq;
            p3 = d2 <= c3sq;
        }
        const unsigned b2 = __ballot_sync(0xffffffffu, p2), b3 = __ballot_sync(0xffffffffu, p3);
        if (p2) {
            const int pos = n2 + __popc(b2 & lt);
            if (pos < ROW_CAP2) nbr2[(size_t)i * ROW_CAP2 + pos] = j;
        }
        if (p3) {
            const int pos = n3 + __popc(b3 & lt);
            if (pos < ROW_CAP3) nbr3[(size_t)i * ROW_CAP3 + pos] = j;
        }
        n2 += __popc(b2);
        n3 += __popc(b3);
    }
    if (lane == 0) {
        if (n2 > ROW_CAP2 || n3 > ROW_CAP3) atomicExch(overflow, 1);
        low2[i] = min(n2, ROW_CAP2);
        low3[i] = min(n3, ROW_CAP3);
        ro2[i] = (int64_t)i * ROW_CAP2;
        ro3[i] = (int64_t)i * ROW_CAP3;
    }
}

// sliced build: flag the 0.45 nm neighbours J < i_lo of the owners [i_lo, i_hi) -- their rows are needed to close the
// chain triplets (I, J, K in nearby[J]) of this rank's owners
__global__ void mark_halo_kernel(int i_lo, int i_hi, const int* __restrict__ low3, const int* __restrict__ nbr3,
                                 unsigned char* __restrict__ halo) {
    const int i = i_lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i_hi) return;
    const int* row = nbr3 + (size_t)i * ROW_CAP3;
    for (int k = 0; k < low3[i]; ++k)
        if (row[k] < i_lo) halo[row[k]] = 1;
}

// own3[i] = low3[i] for owners of this rank, 0 elsewhere (halo rows must not emit (J, K) pairs of their own)
__global__ void mask_range_kernel(int n, int i_lo, int i_hi, const int* __restrict__ low3, int* __restrict__ own3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) own3[i] = (i >= i_lo && i < i_hi) ? low3[i] : 0;
}

// pairs[p] = (I, J) for J in the lower part of I's row; p = pstart[I] + rank
__global__ void emit_pairs_kernel(int n, const int64_t* __restrict__ ro, const int* __restrict__ nbr,
                                  const int* __restrict__ low, const int64_t* __restrict__ pstart,
                                  int32_t* __restrict__ pairs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int* row = nbr + ro[i];
    int32_t* out = pairs + 2 * pstart[i];
    for (int k = 0; k < low[i]; ++k) {
        out[2 * k] = i;
        out[2 * k + 1] = row[k];
    }
}

// number of triplets hanging off each (I,J) pair of the 3B pair list = |nearby[J]|
__global__ void triplet_count_kernel(int64_t nq, const int32_t* __restrict__ qpairs, const int* __restrict__ low3,
                                     int* __restrict__ cnt) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    cnt[q] = low3[qpairs[2 * q + 1]];
}

__global__ void emit_triplets_kernel(int64_t nq, const int32_t* __restrict__ qpairs, const int64_t* __restrict__ ro3,
                                     const int* __restrict__ nbr3, const int* __restrict__ low3,
                                     const int64_t* __restrict__ tstart, int32_t* __restrict__ trips) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int I = qpairs[2 * q], J = qpairs[2 * q + 1];
    const int* row = nbr3 + ro3[J];
    int32_t* out = trips + 3 * tstart[q];
    for (int k = 0; k < low3[J]; ++k) {
        out[3 * k] = I;
        out[3 * k + 1] = J;
        out[3 * k + 2] = row[k];
    }
}

// cumulative work W(I) = 30601 * pairs owned by molecules < I + 111818 * triplets owned by molecules < I
__global__ void owner_weight_kernel(int n, const int64_t* __restrict__ pstart, const int64_t* __restrict__ qstart,
                                    const int64_t* __restrict__ tstart, int has2, int has3, int64_t* __restrict__ W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int64_t w = 0;
    if (has2) w += (int64_t)MBPOL_FLOP_PER_DIMER * pstart[i];
    if (has3) w += (int64_t)MBPOL_FLOP_PER_TRIMER * tstart[qstart[i]];
    W[i] = w;
}

__global__ void contribution_keys_kernel(int64_t first, int64_t count, int width, const int32_t* __restrict__ tuples,
                                         int32_t* __restrict__ keys, int32_t* __restrict__ vals) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= count * width) return;
    const int64_t t = e / width;
    const int r = (int)(e - t * width);
    keys[e] = tuples[width * (first + t) + r];
    vals[e] = (int32_t)e;                       // (local tuple index) * width + role
}

__global__ void segment_offsets_kernel(int n, int64_t count, const int32_t* __restrict__ sorted_keys,
                                       int64_t* __restrict__ off) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m > n) return;
    int64_t lo = 0, hi = count;                 // lower_bound(m)
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (sorted_keys[mid] < m) lo = mid + 1; else hi = mid;
    }
    off[m] = lo;
}

// owner slice of `rank`: boundaries by binary search on the cumulative work, then the matching tuple offsets
__global__ void partition_kernel(int n, int rank, int world, const int64_t* __restrict__ W,
                                 const int64_t* __restrict__ pstart, const int64_t* __restrict__ qstart,
                                 const int64_t* __restrict__ tstart, int64_t* __restrict__ out /* 6 */, int fixed_lo,
                                 int fixed_hi) {
    if (threadIdx.x > 1) return;
    const int r = rank + threadIdx.x;                 // thread 0: lower bound, thread 1: upper bound
    int b;
    if (fixed_hi >= 0) b = threadIdx.x ? fixed_hi : fixed_lo;     // keep the owner range of an earlier balance
    else if (r <= 0) b = 0;
    else if (r >= world) b = n;
    else {
        const double target = (double)W[n] * r / world;
        int lo = 0, hi = n;
        while (lo < hi) {
            const int mid = (lo + hi) / 2;
            if ((double)W[mid] < target) lo = mid + 1; else hi = mid;
        }
        b = lo;
    }
    out[3 * threadIdx.x] = b;
    out[3 * threadIdx.x + 1] = pstart[b];
    out[3 * threadIdx.x + 2] = tstart[qstart[b]];
}

inline unsigned blocks(int64_t n, int b = 128) { return (unsigned)((n + b - 1) / b); }

#define CK(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                          \
            return MBPOL_ECUDA;                                                                \
        }                                                                                      \
    } while (0)

}  // namespace

int DeviceBuf::reserve(size_t bytes, std::string& err) {
    if (bytes <= cap) return MBPOL_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&ptr, want) != cudaSuccess) {
        cudaGetLastError();
        err = "cudaMalloc of " + std::to_string(want) + " bytes failed";
        return MBPOL_ENOMEM;
    }
    cap = want;
    return MBPOL_OK;
}
DeviceBuf::~DeviceBuf() {
    if (ptr) cudaFree(ptr);
}

struct ToInt64 {
    __host__ __device__ int64_t operator()(int v) const { return (int64_t)v; }
};

static int exclusive_scan(const int* in, int64_t* out, int64_t n, DeviceBuf& tmp, cudaStream_t s, std::string& err) {
    // out[0..n] (n+1 entries): out[n] = total (in[n] must be 0).  Input is widened to int64 on the fly.
    cub::TransformInputIterator<int64_t, ToInt64, const int*> it(in, ToInt64());
    size_t bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, out, n + 1, s));
    int rc = tmp.reserve(bytes, err);
    if (rc) return rc;
    CK(cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, it, out, n + 1, s));
    MBPOL_COUNT_LAUNCH(1);
    return MBPOL_OK;
}

int TupleLists::build(const double* pos, const int32_t* idx, int n_mol, const double box[3], bool periodic,
                      double cut2, double cut3, bool want2, bool want3, cudaStream_t s, std::string& err,
                      int slice_lo, int slice_hi) {
    n = n_mol;
    // sliced build (multi-GPU, csrc/capi.cu eval_pass): only the rows this rank needs -- owners [slice_lo, slice_hi)
    // and the 0.45 nm halo below them -- instead of all n; the lists then hold exactly this rank's tuples
    bool slice = slice_hi >= 0 && !force_two_pass && (slice_lo > 0 || slice_hi < n_mol);
    sliced = false;
    rows_reused = false;
    n_pairs = n_trips = n_qpairs = 0;
    has2 = want2;
    has3 = want3;
    if (n <= 0 || (!want2 && !want3)) return MBPOL_OK;
    int rc;
#define RES(buf, count, type) if ((rc = buf.reserve(sizeof(type) * (size_t)(count), err))) return rc
    RES(posO, 3 * n, double);
    gather_oxygens_kernel<<<blocks(n), 128, 0, s>>>(pos, idx, n, posO.as<double>()); MBPOL_COUNT_LAUNCH(1);

    // ---- Verlet skin: can the super rows of an earlier build be filtered instead of scanning the cells again? --------
    const double cmax = fmax(want2 ? cut2 : 0.0, want3 ? cut3 : 0.0);
    const bool use_skin = skin > 0.0 && !slice && !force_two_pass && cmax < 1e9 &&
                          (!periodic || (2.0 * (cmax + skin) <= box[0] && 2.0 * (cmax + skin) <= box[1] &&
                                         2.0 * (cmax + skin) <= box[2]));
    GridArg g;
    g.periodic = periodic ? 1 : 0;
    for (int d = 0; d < 3; ++d) {
        g.box[d] = periodic ? box[d] : 0.0;
        g.inv_box[d] = periodic ? 1.0 / box[d] : 0.0;
        g.lo[d] = 0.0;
        g.inv_extent[d] = 0.0;
        g.nc[d] = 1;
    }
    bool reuse_rows = false;
    if (use_skin && skin_valid && skin_n == n && skin_periodic == periodic && skin_cmax == cmax && skin_built == skin &&
        skin_want2 == want2 && skin_want3 == want3 &&
        (!periodic || (skin_box[0] == box[0] && skin_box[1] == box[1] && skin_box[2] == box[2]))) {
        RES(scratch, 8, int64_t);
        CK(cudaMemsetAsync(scratch.ptr, 0, sizeof(int), s));
        displacement_kernel<<<blocks(n), 128, 0, s>>>(posO.as<double>(), pos_ref.as<double>(), n, g, 0.25 * skin * skin,
                                                     scratch.as<int>()); MBPOL_COUNT_LAUNCH(1);
        int moved = 1;
        CK(cudaMemcpyAsync(&moved, scratch.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        reuse_rows = !moved;
    }
    if (!use_skin) skin_valid = false;
    rows_reused = reuse_rows;

    if (!reuse_rows) {
    // ---- grid ---------------------------------------------------------------------------------------------------
    // rows are only needed out to the largest finite cutoff in use (+ skin); "no cutoff" (1e10) degenerates to one cell
    const double rc_max = (cmax + (use_skin ? skin : 0.0)) * (1.0 + 1e-6);
    double extent[3];
    if (periodic) {
        for (int d = 0; d < 3; ++d) extent[d] = box[d];
    } else {
        RES(scratch, 6, double);
        bbox_kernel<<<1, 256, 0, s>>>(posO.as<double>(), n, scratch.as<double>()); MBPOL_COUNT_LAUNCH(1);
        double bb[6];
        CK(cudaMemcpyAsync(bb, scratch.ptr, sizeof bb, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        for (int d = 0; d < 3; ++d) {
            g.lo[d] = bb[d];
            extent[d] = fmax(bb[3 + d] - bb[d], 1e-9) * (1.0 + 1e-9);
        }
    }
    int64_t ncell = 1;
    for (int d = 0; d < 3; ++d) {
        double q = floor(extent[d] / rc_max);
        g.nc[d] = (int)fmin(fmax(q, 1.0), 1024.0);
        ncell *= g.nc[d];
    }
    while (ncell > 8LL * n + 64) {             // sparse system: coarsen, keeps cell edge >= cutoff
        int d = g.nc[0] >= g.nc[1] && g.nc[0] >= g.nc[2] ? 0 : (g.nc[1] >= g.nc[2] ? 1 : 2);
        ncell /= g.nc[d];
        g.nc[d] = (g.nc[d] + 1) / 2;
        ncell *= g.nc[d];
    }
    for (int d = 0; d < 3; ++d) g.inv_extent[d] = 1.0 / extent[d];

    // ---- cells --------------------------------------------------------------------------------------------------
    ++n_row_builds;
    RES(cell_of, n, int);
    RES(cell_count, ncell + 1, int);
    RES(cell_start64, ncell + 1, int64_t);
    RES(cell_start, ncell + 1, int);
    RES(cell_cursor, ncell, int);
    RES(cell_members, n, int);
    CK(cudaMemsetAsync(cell_count.ptr, 0, sizeof(int) * (ncell + 1), s));
    CK(cudaMemsetAsync(cell_cursor.ptr, 0, sizeof(int) * ncell, s));
    cell_count_kernel<<<blocks(n), 128, 0, s>>>(posO.as<double>(), n, g, cell_of.as<int>(), cell_count.as<int>()); MBPOL_COUNT_LAUNCH(1);
    {
        size_t bytes = 0;
        CK(cub::DeviceScan::ExclusiveSum(nullptr, bytes, cell_count.as<int>(), cell_start.as<int>(), (int)ncell + 1, s));
        if ((rc = cub_tmp.reserve(bytes, err))) return rc;
        CK(cub::DeviceScan::ExclusiveSum(cub_tmp.ptr, bytes, cell_count.as<int>(), cell_start.as<int>(), (int)ncell + 1, s));
        MBPOL_COUNT_LAUNCH(1);
    }
    cell_fill_kernel<<<blocks(n), 128, 0, s>>>(n, cell_of.as<int>(), cell_start.as<int>(), cell_cursor.as<int>(),
                                               cell_members.as<int>()); MBPOL_COUNT_LAUNCH(1);
    }

    // ---- neighbour rows -----------------------------------------------------------------------------------------
    const double c2sq = want2 ? cut2 * cut2 : -1.0, c3sq = want3 ? cut3 * cut3 : -1.0;
    RES(cnt2, n + 1, int);
    RES(cnt3, n + 1, int);
    RES(ro2, n + 1, int64_t);
    RES(ro3, n + 1, int64_t);
    RES(low2, n + 1, int);
    RES(low3, n + 1, int);
    int64_t tot[2];
    auto two_pass_rows = [&]() -> int {                 // general CSR path: any row length
        CK(cudaMemsetAsync(cnt2.ptr, 0, sizeof(int) * (n + 1), s));
        CK(cudaMemsetAsync(cnt3.ptr, 0, sizeof(int) * (n + 1), s));
        CK(cudaMemsetAsync(low2.ptr, 0, sizeof(int) * (n + 1), s));
        CK(cudaMemsetAsync(low3.ptr, 0, sizeof(int) * (n + 1), s));
        row_count_kernel<<<blocks(n), 128, 0, s>>>(posO.as<double>(), n, g, cell_start.as<int>(),
                                                   cell_members.as<int>(), c2sq, c3sq, cnt2.as<int>(), cnt3.as<int>());
        MBPOL_COUNT_LAUNCH(1);
        int r;
        if ((r = exclusive_scan(cnt2.as<int>(), ro2.as<int64_t>(), n, cub_tmp, s, err))) return r;
        if ((r = exclusive_scan(cnt3.as<int>(), ro3.as<int64_t>(), n, cub_tmp, s, err))) return r;
        CK(cudaMemcpyAsync(&tot[0], ro2.as<int64_t>() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&tot[1], ro3.as<int64_t>() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if ((r = nbr2.reserve(sizeof(int) * (size_t)(tot[0] + 1), err))) return r;
        if ((r = nbr3.reserve(sizeof(int) * (size_t)(tot[1] + 1), err))) return r;
        row_fill_kernel<<<blocks(n), 128, 0, s>>>(posO.as<double>(), n, g, cell_start.as<int>(), cell_members.as<int>(),
                                                  c2sq, c3sq, ro2.as<int64_t>(), ro3.as<int64_t>(), nbr2.as<int>(),
                                                  nbr3.as<int>(), low2.as<int>(), low3.as<int>());
        MBPOL_COUNT_LAUNCH(1);
        return MBPOL_OK;
    };
    const int* q_count = nullptr;                        // per-owner length of the 3B pair rows that are emitted
    auto pair_offsets = [&](int* overflow_host) -> int {
        int r;
        if ((r = exclusive_scan(low2.as<int>(), pstart.as<int64_t>(), n, cub_tmp, s, err))) return r;
        if ((r = exclusive_scan(q_count, qstart.as<int64_t>(), n, cub_tmp, s, err))) return r;
        CK(cudaMemcpyAsync(&tot[0], pstart.as<int64_t>() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(&tot[1], qstart.as<int64_t>() + n, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        if (overflow_host) CK(cudaMemcpyAsync(overflow_host, scratch.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return MBPOL_OK;
    };
    RES(pstart, n + 1, int64_t);
    RES(qstart, n + 1, int64_t);
    bool fast = !force_two_pass;
    if (fast) {
        RES(nbr2, (size_t)n * ROW_CAP2 + 1, int);
        RES(nbr3, (size_t)n * ROW_CAP3 + 1, int);
        RES(scratch, 8, int64_t);
        CK(cudaMemsetAsync(scratch.ptr, 0, sizeof(int), s));
        CK(cudaMemsetAsync(low2.ptr, 0, sizeof(int) * (n + 1), s));
        CK(cudaMemsetAsync(low3.ptr, 0, sizeof(int) * (n + 1), s));
        if (use_skin) {
            if (!reuse_rows) {                           // super rows out to cmax + skin, positions remembered
                RES(snbr, (size_t)n * SKIN_CAP + 1, int);
                RES(slow, n + 1, int);
                RES(sk_low3, n + 1, int);
                RES(sk_ro, 2 * (size_t)(n + 1), int64_t);
                RES(pos_ref, 3 * n, double);
                CK(cudaMemsetAsync(slow.ptr, 0, sizeof(int) * (n + 1), s));
                const double cs = (cmax + skin) * (cmax + skin);
                rows_warp_kernel<SKIN_CAP><<<blocks(n, ROWS_WARPS), ROWS_WARPS * 32, 0, s>>>(
                    posO.as<double>(), n, g, cell_start.as<int>(), cell_members.as<int>(), cs, -1.0, snbr.as<int>(),
                    nbr3.as<int>(), slow.as<int>(), sk_low3.as<int>(), sk_ro.as<int64_t>(), sk_ro.as<int64_t>() + n + 1,
                    scratch.as<int>(), 0, n, nullptr);
                CK(cudaMemcpyAsync(pos_ref.ptr, posO.ptr, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, s));
                MBPOL_COUNT_LAUNCH(1);
            }
            filter_rows_kernel<<<blocks(n, ROWS_WARPS), ROWS_WARPS * 32, 0, s>>>(
                posO.as<double>(), n, g, snbr.as<int>(), slow.as<int>(), c2sq, c3sq, nbr2.as<int>(), nbr3.as<int>(),
                low2.as<int>(), low3.as<int>(), ro2.as<int64_t>(), ro3.as<int64_t>(), scratch.as<int>());
            MBPOL_COUNT_LAUNCH(1);
        } else {
        rows_warp_kernel<ROW_CAP2><<<blocks(n, ROWS_WARPS), ROWS_WARPS * 32, 0, s>>>(
            posO.as<double>(), n, g, cell_start.as<int>(), cell_members.as<int>(), c2sq, c3sq, nbr2.as<int>(),
            nbr3.as<int>(), low2.as<int>(), low3.as<int>(), ro2.as<int64_t>(), ro3.as<int64_t>(), scratch.as<int>(),
            slice ? slice_lo : 0, slice ? slice_hi : n, nullptr);
        MBPOL_COUNT_LAUNCH(1);
        }
        q_count = low3.as<int>();
        if (slice && want3) {
            RES(halo, n + 1, unsigned char);
            RES(own3, n + 1, int);
            CK(cudaMemsetAsync(halo.ptr, 0, (size_t)n + 1, s));
            if (slice_hi > slice_lo)                     // (an empty owner range has no halo)
                mark_halo_kernel<<<blocks(slice_hi - slice_lo), 128, 0, s>>>(slice_lo, slice_hi, low3.as<int>(),
                                                                             nbr3.as<int>(), halo.as<unsigned char>());
            rows_warp_kernel<ROW_CAP2><<<blocks(n, ROWS_WARPS), ROWS_WARPS * 32, 0, s>>>(
                posO.as<double>(), n, g, cell_start.as<int>(), cell_members.as<int>(), -1.0, c3sq, nbr2.as<int>(),
                nbr3.as<int>(), low2.as<int>(), low3.as<int>(), ro2.as<int64_t>(), ro3.as<int64_t>(),
                scratch.as<int>(), slice_lo, slice_hi, halo.as<unsigned char>());
            mask_range_kernel<<<blocks(n + 1), 128, 0, s>>>(n, slice_lo, slice_hi, low3.as<int>(), own3.as<int>());
            MBPOL_COUNT_LAUNCH(3);
            q_count = own3.as<int>();
        }
        int overflow = 0;
        if ((rc = pair_offsets(&overflow))) return rc;
        if (overflow) fast = false;                      // some row exceeded the fixed stride: redo with the CSR path
        if (use_skin) {
            skin_valid = !overflow;
            skin_n = n;
            skin_periodic = periodic;
            skin_cmax = cmax;
            skin_built = skin;
            skin_want2 = want2;
            skin_want3 = want3;
            for (int d = 0; d < 3; ++d) skin_box[d] = box[d];
        }
    }
    if (!fast && reuse_rows) {                           // a filtered 3B row overflowed its stride: the general path needs
        skin_valid = false;                              // the cell grid, which a reused build does not have -> start over
        return build(pos, idx, n_mol, box, periodic, cut2, cut3, want2, want3, s, err, slice_lo, slice_hi);
    }
    if (!fast) {
        slice = false;                                   // the general CSR path always builds every row
        if ((rc = two_pass_rows())) return rc;
        q_count = low3.as<int>();
        if ((rc = pair_offsets(nullptr))) return rc;
    }
    used_fast_rows = fast;
    n_pairs = tot[0];
    n_qpairs = tot[1];
    RES(pairs, 2 * n_pairs + 2, int32_t);
    RES(qpairs, 2 * n_qpairs + 2, int32_t);
    emit_pairs_kernel<<<blocks(n), 128, 0, s>>>(n, ro2.as<int64_t>(), nbr2.as<int>(), low2.as<int>(),
                                                pstart.as<int64_t>(), pairs.as<int32_t>()); MBPOL_COUNT_LAUNCH(1);
    emit_pairs_kernel<<<blocks(n), 128, 0, s>>>(n, ro3.as<int64_t>(), nbr3.as<int>(), q_count,
                                                qstart.as<int64_t>(), qpairs.as<int32_t>()); MBPOL_COUNT_LAUNCH(1);

    // ---- triplets -------------------------------------------------------------------------------------------------
    RES(tcnt, n_qpairs + 1, int);
    RES(tstart, n_qpairs + 1, int64_t);
    CK(cudaMemsetAsync(tcnt.ptr, 0, sizeof(int) * (n_qpairs + 1), s));
    if (n_qpairs > 0)
        triplet_count_kernel<<<blocks(n_qpairs), 128, 0, s>>>(n_qpairs, qpairs.as<int32_t>(), low3.as<int>(),
                                                              tcnt.as<int>()); MBPOL_COUNT_LAUNCH(1);
    if ((rc = exclusive_scan(tcnt.as<int>(), tstart.as<int64_t>(), n_qpairs, cub_tmp, s, err))) return rc;
    CK(cudaMemcpyAsync(&tot[0], tstart.as<int64_t>() + n_qpairs, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    n_trips = tot[0];
    RES(trips, 3 * n_trips + 3, int32_t);
    if (n_qpairs > 0)
        emit_triplets_kernel<<<blocks(n_qpairs), 128, 0, s>>>(n_qpairs, qpairs.as<int32_t>(), ro3.as<int64_t>(),
                                                              nbr3.as<int>(), low3.as<int>(), tstart.as<int64_t>(),
                                                              trips.as<int32_t>()); MBPOL_COUNT_LAUNCH(1);
    if (slice) {
        // the lists hold this rank's tuples only; the slice boundaries stay those of the last full build
        sliced = true;
        own_i0 = slice_lo;
        own_i1 = slice_hi;
        own_p0 = own_t0 = 0;
        own_p1 = n_pairs;
        own_t1 = n_trips;
        CK(cudaGetLastError());
        return MBPOL_OK;
    }
    // ---- cumulative owner weights for the multi-GPU partition -------------------------------------------------------
    RES(weight, n + 1, int64_t);
    owner_weight_kernel<<<blocks(n + 1), 128, 0, s>>>(n, pstart.as<int64_t>(), qstart.as<int64_t>(),
                                                      tstart.as<int64_t>(), want2, want3, weight.as<int64_t>()); MBPOL_COUNT_LAUNCH(1);
    CK(cudaGetLastError());
#undef RES
    return MBPOL_OK;
}

int TupleLists::partition(int rank, int world, cudaStream_t s, std::string& err, int fixed_lo, int fixed_hi) {
    // Owner ranges [I0, I1) with equal cumulative work; every rank computes the same boundaries from the same lists.
    // With fixed_hi >= 0 the range [fixed_lo, fixed_hi) of an earlier balance is kept and only its pair / triplet
    // offsets in the (full) lists of this build are looked up.
    own_p0 = own_t0 = 0;
    own_p1 = n_pairs;
    own_t1 = n_trips;
    own_i0 = 0;
    own_i1 = n;
    if (world <= 1 || n <= 0) return MBPOL_OK;
    int rc;
    if ((rc = scratch.reserve(sizeof(int64_t) * 6, err))) return rc;
    partition_kernel<<<1, 32, 0, s>>>(n, rank, world, weight.as<int64_t>(), pstart.as<int64_t>(), qstart.as<int64_t>(),
                                      tstart.as<int64_t>(), scratch.as<int64_t>(), fixed_lo, fixed_hi); MBPOL_COUNT_LAUNCH(1);
    int64_t h[6];
    CK(cudaMemcpyAsync(h, scratch.ptr, sizeof h, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    own_i0 = (int)h[0];
    own_p0 = h[1];
    own_t0 = h[2];
    own_i1 = (int)h[3];
    own_p1 = h[4];
    own_t1 = h[5];
    return MBPOL_OK;
}

int ContributionIndex::build(const int32_t* tuples, int width, int64_t first, int64_t count, int n_mol,
                             DeviceBuf& cub_tmp, cudaStream_t s, std::string& err) {
    n_entries = count * width;
    int rc;
    if ((rc = off.reserve(sizeof(int64_t) * (n_mol + 1), err))) return rc;
    if (n_entries == 0) {
        CK(cudaMemsetAsync(off.ptr, 0, sizeof(int64_t) * (n_mol + 1), s));
        return MBPOL_OK;
    }
    if (n_entries > 0x7FFFFFFFLL) {
        err = "too many tuple contributions for 32-bit entry ids";
        return MBPOL_EINVAL;
    }
    if ((rc = keys_in.reserve(sizeof(int32_t) * n_entries, err))) return rc;
    if ((rc = vals_in.reserve(sizeof(int32_t) * n_entries, err))) return rc;
    if ((rc = keys.reserve(sizeof(int32_t) * n_entries, err))) return rc;
    if ((rc = vals.reserve(sizeof(int32_t) * n_entries, err))) return rc;
    contribution_keys_kernel<<<blocks(n_entries), 128, 0, s>>>(first, count, width, tuples, keys_in.as<int32_t>(),
                                                               vals_in.as<int32_t>()); MBPOL_COUNT_LAUNCH(1);
    int bits = 1;
    while ((1LL << bits) < n_mol) ++bits;
    size_t bytes = 0;
    // LSD radix sort is stable: inside one molecule the entries stay in ascending (tuple, role) order -> the
    // summation order of the force reduction is fixed by the canonical tuple order alone.
    CK(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in.as<int32_t>(), keys.as<int32_t>(),
                                       vals_in.as<int32_t>(), vals.as<int32_t>(), (int)n_entries, 0, bits, s));
    if ((rc = cub_tmp.reserve(bytes, err))) return rc;
    CK(cub::DeviceRadixSort::SortPairs(cub_tmp.ptr, bytes, keys_in.as<int32_t>(), keys.as<int32_t>(),
                                       vals_in.as<int32_t>(), vals.as<int32_t>(), (int)n_entries, 0, bits, s));
    MBPOL_COUNT_LAUNCH(1);
    segment_offsets_kernel<<<blocks(n_mol + 1), 128, 0, s>>>(n_mol, n_entries, keys.as<int32_t>(), off.as<int64_t>()); MBPOL_COUNT_LAUNCH(1);
    CK(cudaGetLastError());
    return MBPOL_OK;
}

}  // namespace mbpol
