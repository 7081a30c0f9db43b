// Device-side tuple lists (lists.cu) and the tuple->molecule contribution index used by reduce.cu.
#pragma once

#include <string>
#include <vector>

#include "internal.h"

namespace mbpol {

struct GridArg {
    double box[3];         // periodic box edges (nm)
    double lo[3];          // non-periodic: lower corner of the bounding box
    double inv_extent[3];
    double inv_box[3];     // screening only (pair_d2_fast)
    int nc[3];             // cells per axis (cell edge >= largest cutoff)
    int periodic;
};

// grow-only device allocation
struct DeviceBuf {
    void* ptr = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes, std::string& err);
    template <class T>
    T* as() const { return static_cast<T*>(ptr); }
    DeviceBuf() = default;
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
    ~DeviceBuf();
};

struct TupleLists {
    int n = 0;                       // molecules
    bool has2 = false, has3 = false;
    bool force_two_pass = false;     // MBPOL_FORCE_TWO_PASS=1: skip the fixed-stride fast path (tests cover both)
    bool used_fast_rows = false;
    bool sliced = false;             // the last build covered only this rank's owners [own_i0, own_i1) + their halo
    bool rows_reused = false;        // the last build filtered the skin rows of an earlier build instead of scanning cells
    int64_t n_row_builds = 0;        // cell-grid + neighbour-row scans so far
    int64_t n_pairs = 0, n_qpairs = 0, n_trips = 0;
    // owner partition of this rank: molecules [own_i0, own_i1), pairs [own_p0, own_p1), triplets [own_t0, own_t1)
    int own_i0 = 0, own_i1 = 0;
    int64_t own_p0 = 0, own_p1 = 0, own_t0 = 0, own_t1 = 0;

    DeviceBuf posO, scratch, cub_tmp;
    DeviceBuf cell_of, cell_count, cell_start64, cell_start, cell_cursor, cell_members;
    DeviceBuf cnt2, cnt3, ro2, ro3, low2, low3, nbr2, nbr3;
    DeviceBuf pstart, qstart, pairs, qpairs, tcnt, tstart, trips, weight, halo, own3;
    // Verlet skin (0 = rebuild the rows on every call, the reference's behaviour): super rows + the positions they were
    // built from, and what they were built for
    double skin = 0.0;
    bool skin_valid = false, skin_periodic = false, skin_want2 = false, skin_want3 = false;
    int skin_n = 0;
    double skin_cmax = 0.0, skin_built = 0.0, skin_box[3] = {0, 0, 0};
    DeviceBuf snbr, slow, sk_low3, sk_ro, pos_ref;

    int build(const double* pos, const int32_t* idx, int n_mol, const double box[3], bool periodic, double cut2,
              double cut3, bool want2, bool want3, cudaStream_t s, std::string& err, int slice_lo = 0,
              int slice_hi = -1);
    int partition(int rank, int world, cudaStream_t s, std::string& err, int fixed_lo = -1, int fixed_hi = -1);
};

// For every molecule m: the list of (local tuple, role) entries that put force on it, in ascending (tuple, role)
// order; entries of m are vals[off[m] .. off[m+1]).
struct ContributionIndex {
    int64_t n_entries = 0;
    DeviceBuf keys_in, vals_in, keys, vals, off;
    int build(const int32_t* tuples, int width, int64_t first, int64_t count, int n_mol, DeviceBuf& cub_tmp,
              cudaStream_t s, std::string& err);
};

// reduce.cu -----------------------------------------------------------------------------------------------------------
// f_mol[m][9] (+)= sum over the contributions of m of the 3x3 block of its role, fixed order; one warp per molecule.
void launch_force_reduce(const ContributionIndex& ci, int width, const double* f_tuple, int n_mol, double* f_mol,
                         bool accumulate, cudaStream_t s);
// force[idx[m][k]][c] += f_mol[m][3k+c]
void launch_force_scatter(const double* f_mol, const int32_t* idx, int n_mol, double* force, cudaStream_t s);
// fixed-point variant for OpenMM's force buffer: buf[c*padded + particle] += llround(f * 2^32)
void launch_force_scatter_fixed(const double* f_mol, const int32_t* idx, int n_mol, long long* buf, int padded,
                                cudaStream_t s);
// deterministic sum of e[0..n) into *out (device); two fixed-shape passes
void launch_energy_sum(const double* e, int64_t n, double* partial /* >= 1024 doubles */, double* out, cudaStream_t s);

}  // namespace mbpol
