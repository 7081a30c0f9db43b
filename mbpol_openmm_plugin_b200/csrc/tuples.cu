// Tuple stage of the MB-pol 2B/3B path: everything around the polynomial.
//
//   *_vars   : gather O/H/H, minimum-image the molecules, build the 31 (2B) / 36 (3B) PIP variables x_i
//   *_forces : chain rule from x_i dE/dx_i back to the atoms, switching function, per-tuple energy + force block
//
// One thread per tuple; variables are written tuple-fastest ([var][tuple]) so that the generated polynomial kernel
// (gen/pip{2b,3b}_eval.ptx) reads and writes fully coalesced.  The per-tuple force blocks are reduced afterwards by
// the deterministic segmented reduction in reduce.cu -- no atomics anywhere.
//
// HBM traffic (round 2).  Every variable is exp(k (r0 - r)) [/ r] of ONE distance, so d x_i / d r = -(k [+ 1/r]) x_i and
// the chain rule needs only the product x_i dE/dx_i -- which the polynomial kernel writes directly (mode bit 1 of
// pip{2b,3b}_eval) -- and 1/r, which the force kernel recomputes from the coordinates it gathers anyway (one rsqrt, no
// exp).  The force kernels therefore read the polynomial output only; x is read once (by the polynomial kernel) and
// 1/r never touches memory: 1,392 B per trimer through HBM instead of 2,256 B (2B: 1,050 instead of 1,796).
// The sites live in a shared-memory row per thread during the chain rule, not in registers, and the variables are
// staged through shared memory with cp.async so that all its loads are in flight at once (see stage_poly).
//
// Behaviour follows the Reference platform:
//   2B  MBPolReferenceTwoBodyForce::calculatePairIxn      platforms/reference/src/MBPolReferenceTwoBodyForce.cpp:110-296
//       variable::v_exp / v_coul / grads, monomer::setup / grads, f_switch
//                                                         platforms/reference/src/mbpol_2body_constants.cpp:10-111
//   3B  MBPolReferenceThreeBodyForce::calculateTripletIxn platforms/reference/src/MBPolReferenceThreeBodyForce.cpp:122-293
//       var / g_var / threebody_f_switch                  platforms/reference/src/MBPolReferenceThreeBodyForce.cpp:67-120
//   PBC imageParticles / imageMolecules                   platforms/reference/src/MBPolReferenceTwoBodyForce.cpp:66-109
#include <cuda_pipeline_primitives.h>

#include "internal.h"

namespace mbpol {

namespace {

struct V3 {
    double x, y, z;
};
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ void operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ void operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; }
__device__ __forceinline__ double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

__device__ __forceinline__ V3 load3(const double* __restrict__ pos, int p) {
    const double* q = pos + 3 * (size_t)p;
    return {__ldg(q), __ldg(q + 1), __ldg(q + 2)};
}

// p += box * floor((ref - p)/box + 0.5), per axis
__device__ __forceinline__ void image_to(const double bx, const double by, const double bz, const V3& ref, V3& p) {
    p.x += bx * floor((ref.x - p.x) / bx + 0.5);
    p.y += by * floor((ref.y - p.y) / by + 0.5);
    p.z += bz * floor((ref.z - p.z) / bz + 0.5);
}

// Load one water (O,H1,H2) scaled by `scale`, hydrogens imaged to their own oxygen.
__device__ __forceinline__ void load_water(const double* __restrict__ pos, const int32_t* __restrict__ idx, int mol,
                                           double scale, V3& O, V3& H1, V3& H2) {
    O = load3(pos, __ldg(idx + 3 * mol)) * scale;
    H1 = load3(pos, __ldg(idx + 3 * mol + 1)) * scale;
    H2 = load3(pos, __ldg(idx + 3 * mol + 2)) * scale;
}

// cosine switch on [ri, rf]: value and d/dr
__device__ __forceinline__ double f_switch(double r, double ri, double rf, double& g) {
    if (r > rf) {
        g = 0.0;
        return 0.0;
    } else if (r > ri) {
        const double t1 = M_PI / (rf - ri);
        const double x = (r - ri) * t1;
        double s, c;
        sincos(x, &s, &c);
        g = -s * t1 / 2.0;
        return (1.0 + c) / 2.0;
    }
    g = 0.0;
    return 1.0;
}

// Per-tuple force blocks are W = 18 / 27 consecutive doubles: written straight from registers every lane would touch its
// own 32-byte sectors (stride 144 / 216 B).  Stage the CTA's blocks in shared memory and stream them out contiguously.
template <int W, int NT>
__device__ __forceinline__ void store_blocks(double* __restrict__ f_out, int64_t cta_first, int64_t n, const double* sh) {
    __syncthreads();
    const int64_t count = min((int64_t)NT, n - cta_first) * W;
    double* dst = f_out + cta_first * W;
    for (int64_t i = threadIdx.x; i < count; i += NT) dst[i] = sh[i];
}

// ---------------------------------------------------------------------------------------------------------------
// 2B.  Sites: 0 Oa, 1 Ha1, 2 Ha2, 3 Ob, 4 Hb1, 5 Hb2, 6 Xa1, 7 Xa2, 8 Xb1, 9 Xb2 (Angstrom).
// Variable table = MBPolReferenceTwoBodyForce.cpp:170-207.  kind 0: exp(k(r0-r)), 1: exp(k(r0-r))/r
#define MBPOL_2B_VARS_A(V)                                                                                    \
    V(0, 0, K2_HH_INTRA, D0_INTRA, 1, 2)  V(1, 0, K2_HH_INTRA, D0_INTRA, 4, 5)                                  \
    V(2, 0, K2_OH_INTRA, D0_INTRA, 0, 1)  V(3, 0, K2_OH_INTRA, D0_INTRA, 0, 2)                                  \
    V(4, 0, K2_OH_INTRA, D0_INTRA, 3, 4)  V(5, 0, K2_OH_INTRA, D0_INTRA, 3, 5)                                  \
    V(6, 1, K2_HH_COUL, D0_INTER, 1, 4)   V(7, 1, K2_HH_COUL, D0_INTER, 1, 5)                                   \
    V(8, 1, K2_HH_COUL, D0_INTER, 2, 4)   V(9, 1, K2_HH_COUL, D0_INTER, 2, 5)
#define MBPOL_2B_VARS_B(V)                                                                                    \
    V(10, 1, K2_OH_COUL, D0_INTER, 0, 4)  V(11, 1, K2_OH_COUL, D0_INTER, 0, 5)                                  \
    V(12, 1, K2_OH_COUL, D0_INTER, 3, 1)  V(13, 1, K2_OH_COUL, D0_INTER, 3, 2)                                  \
    V(14, 1, K2_OO_COUL, D0_INTER, 0, 3)                                                                        \
    V(15, 0, K2_XH_MAIN, D0_INTER, 6, 4)  V(16, 0, K2_XH_MAIN, D0_INTER, 6, 5)                                  \
    V(17, 0, K2_XH_MAIN, D0_INTER, 7, 4)  V(18, 0, K2_XH_MAIN, D0_INTER, 7, 5)                                  \
    V(19, 0, K2_XH_MAIN, D0_INTER, 8, 1)  V(20, 0, K2_XH_MAIN, D0_INTER, 8, 2)
#define MBPOL_2B_VARS_C(V)                                                                                    \
    V(21, 0, K2_XH_MAIN, D0_INTER, 9, 1)  V(22, 0, K2_XH_MAIN, D0_INTER, 9, 2)                                  \
    V(23, 0, K2_XO_MAIN, D0_INTER, 0, 8)  V(24, 0, K2_XO_MAIN, D0_INTER, 0, 9)                                  \
    V(25, 0, K2_XO_MAIN, D0_INTER, 3, 6)  V(26, 0, K2_XO_MAIN, D0_INTER, 3, 7)                                  \
    V(27, 0, K2_XX_MAIN, D0_INTER, 6, 8)  V(28, 0, K2_XX_MAIN, D0_INTER, 6, 9)                                  \
    V(29, 0, K2_XX_MAIN, D0_INTER, 7, 8)  V(30, 0, K2_XX_MAIN, D0_INTER, 7, 9)
#define MBPOL_2B_VARS(V) MBPOL_2B_VARS_A(V) MBPOL_2B_VARS_B(V) MBPOL_2B_VARS_C(V)

// Gather + image one dimer (positions in Angstrom) and place the four lone-pair sites.  Returns false if the
// reference would early-out (rOO > 6.5 A or < 2 A, MBPolReferenceTwoBodyForce.cpp:141-145).
__device__ __forceinline__ bool dimer_sites(const double* __restrict__ pos, const int32_t* __restrict__ idx, int ma,
                                            int mb, const BoxArg& box, V3 (&s)[10], V3& dOO, double& rOO) {
    load_water(pos, idx, ma, NM_TO_A, s[0], s[1], s[2]);
    load_water(pos, idx, mb, NM_TO_A, s[3], s[4], s[5]);
    if (box.periodic) {
        const double bx = box.len[0] * NM_TO_A, by = box.len[1] * NM_TO_A, bz = box.len[2] * NM_TO_A;
        image_to(bx, by, bz, s[0], s[1]);
        image_to(bx, by, bz, s[0], s[2]);
        image_to(bx, by, bz, s[0], s[3]);
        image_to(bx, by, bz, s[3], s[4]);
        image_to(bx, by, bz, s[3], s[5]);
    }
    dOO = s[0] - s[3];
    rOO = sqrt(dot(dOO, dOO));
    if (rOO > R2F || rOO < 2.0) return false;
    // monomer::setup (mbpol_2body_constants.cpp:59-72)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
        const V3 O = s[3 * m], oh1 = s[3 * m + 1] - O, oh2 = s[3 * m + 2] - O;
        const V3 v = cross(oh1, oh2);
        const V3 in_plane = O + (oh1 + oh2) * (0.5 * GAMMA_IN_PLANE);
        const V3 oop = v * GAMMA_OUT_OF_PLANE;
        s[6 + 2 * m] = in_plane + oop;
        s[7 + 2 * m] = in_plane - oop;
    }
    return true;
}

__global__ void __launch_bounds__(128) dimer_vars_kernel(const double* __restrict__ pos, const int32_t* __restrict__ idx,
                                                         const int32_t* __restrict__ pairs, int64_t first, int64_t n,
                                                         int64_t ld, BoxArg box, double* __restrict__ x) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int ma = pairs[2 * (first + t)], mb = pairs[2 * (first + t) + 1];
    V3 s[10], dOO;
    double rOO;
    const bool active = dimer_sites(pos, idx, ma, mb, box, s, dOO, rOO);
#define V(i, kind, k, r0, p, q)                                         \
    {                                                                   \
        double xi = 0.0;                                                \
        if (active) {                                                   \
            const V3 d = s[p] - s[q];                                   \
            const double d2 = dot(d, d);                                \
            const double ri = rsqrt(d2); /* one MUFU + Newton steps */  \
            const double r = d2 * ri;    /* for r and 1/r (<= 2 ulp) */ \
            const double e = exp((k) * ((r0) - r));                     \
            xi = (kind) ? e * ri : e;                                   \
        }                                                               \
        x[(size_t)(i) * ld + t] = xi;                                   \
    }
    MBPOL_2B_VARS(V)
#undef V
}

// The force kernels are latency-bound on the polynomial outputs they read (ncu: 12-14 long-scoreboard stall cycles per
// issue, DRAM 13 %, 12 warps per SM): every thread therefore issues ALL its 1 + NV loads up front as 8-byte cp.async
// copies into a shared-memory column block [1 + NV][128] (coalesced across the CTA, no registers held), does its gather /
// imaging / switching-function work while they fly, and waits once.  Each thread reads back only what it copied itself.
template <int NVARS>
__device__ __forceinline__ void stage_poly(const double* __restrict__ poly, int64_t ld, int64_t t, double* col) {
#pragma unroll
    for (int i = 0; i <= NVARS; ++i) __pipeline_memcpy_async(col + i * 128, poly + (size_t)i * ld + t, sizeof(double));
    __pipeline_commit();
}

// Shared-memory row of one thread: its sites during the chain rule, its force block afterwards.
template <int NS>
struct SiteRow {
    double* p;
    __device__ __forceinline__ V3 get(int i) const { return {p[3 * i], p[3 * i + 1], p[3 * i + 2]}; }
    __device__ __forceinline__ void put(int i, const V3& v) {
        p[3 * i] = v.x;
        p[3 * i + 1] = v.y;
        p[3 * i + 2] = v.z;
    }
};

__global__ void __launch_bounds__(128, 3) dimer_forces_kernel(const double* __restrict__ pos,
                                                              const int32_t* __restrict__ idx,
                                                              const int32_t* __restrict__ pairs, int64_t first, int64_t n,
                                                              int64_t ld, BoxArg box,
                                                              const double* __restrict__ poly /* E, x_i dE/dx_i */,
                                                              double* __restrict__ e_out, double* __restrict__ f_out) {
    extern __shared__ double dyn_smem[];
    double* sh = dyn_smem;                              // [128][30]: 10 sites per thread (stride 30 doubles: conflict-free)
    double* col = dyn_smem + 128 * 30 + threadIdx.x;    // [1 + NV2][128]: this thread's polynomial outputs
    const int64_t cta_first = blockIdx.x * (int64_t)blockDim.x;
    const int64_t t = cta_first + threadIdx.x;
    if (t < n) stage_poly<NV2>(poly, ld, t, col);
    SiteRow<10> row{sh + 30 * threadIdx.x};
    V3 F[10];
#pragma unroll
    for (int c = 0; c < 10; ++c) F[c] = {0.0, 0.0, 0.0};
    V3 dOO;
    double rOO = 0.0, e = 0.0;
    bool active = false;
    if (t < n) {
        V3 s[10];
        active = dimer_sites(pos, idx, pairs[2 * (first + t)], pairs[2 * (first + t) + 1], box, s, dOO, rOO);
#pragma unroll
        for (int c = 0; c < 10; ++c) row.put(c, s[c]);
    }
    __pipeline_wait_prior(0);
    if (active) {
        const double E_poly = col[0];
        // switching function first: its gradient seeds F[Oa], F[Ob], and every later term carries sw * unit conversion,
        // so that nothing but F stays live across the variable batches
        double gsw;
        const double sw = f_switch(rOO, R2I, R2F, gsw);
        const double scale = sw * CAL2JOULE * -10.0;     // gradient (kcal/mol/A) -> force (kJ/mol/nm)
        gsw *= E_poly / rOO;
        const V3 dsw = dOO * (gsw * CAL2JOULE * -10.0);
        F[0] = dsw;
        F[3] = dsw * -1.0;
        e = sw * E_poly * CAL2JOULE;
        // variable::grads with the analytic d(var)/d(r) of v_exp / v_coul (mbpol_2body_constants.cpp:10-55):
        //   d x / d r = -k x (v_exp),  -(k + 1/r) x (v_coul);   xg = x dE/dx comes from the polynomial kernel
#define V(i, kind, k, r0, p, q)                                                      \
    {                                                                                \
        const V3 d = row.get(p) - row.get(q);                                        \
        const double ri = rsqrt(dot(d, d));                                          \
        const double xg = col[((i) + 1) * 128] * scale;                              \
        const double gg = ((kind) ? -((k) + ri) * ri : -(k) * ri) * xg;              \
        const V3 dg = d * gg;                                                        \
        F[p] += dg;                                                                  \
        F[q] -= dg;                                                                  \
    }
        MBPOL_2B_VARS(V)
#undef V
        // monomer::grads: fold the lone-pair gradients back onto O,H,H (mbpol_2body_constants.cpp:76-95); linear in F,
        // so the scale factor folded into gg above passes through
#pragma unroll
        for (int m = 0; m < 2; ++m) {
            const V3 O = row.get(3 * m), oh1 = row.get(3 * m + 1) - O, oh2 = row.get(3 * m + 2) - O;
            const V3 g1 = F[6 + 2 * m], g2 = F[7 + 2 * m];
            const V3 gm = g1 - g2;
            const V3 t1 = cross(oh2, gm), t2 = cross(oh1, gm);
            const V3 gsum = g1 + g2;
            const V3 in_plane = gsum * (0.5 * GAMMA_IN_PLANE);
            const V3 gh1 = in_plane + t1 * GAMMA_OUT_OF_PLANE;
            const V3 gh2 = in_plane - t2 * GAMMA_OUT_OF_PLANE;
            F[3 * m] += gsum - (gh1 + gh2);
            F[3 * m + 1] += gh1;
            F[3 * m + 2] += gh2;
        }
    }
    if (t < n) e_out[t] = e;
    __syncthreads();                                    // every thread is done with its sites: the rows become force blocks
    double* f = sh + 18 * threadIdx.x;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        f[3 * a] = F[a].x;
        f[3 * a + 1] = F[a].y;
        f[3 * a + 2] = F[a].z;
    }
    store_blocks<18, 128>(f_out, cta_first, n, sh);
}

// ---------------------------------------------------------------------------------------------------------------
// 3B.  Sites: 0 Oa, 1 Ha1, 2 Ha2, 3 Ob, 4 Hb1, 5 Hb2, 6 Oc, 7 Hc1, 8 Hc2 (nm; distances are scaled to Angstrom).
// Variable table = MBPolReferenceThreeBodyForce.cpp:170-206: x = exp(-k (d - d0)).
#define MBPOL_3B_VARS_A(V)                                                                                        \
    V(0, K3_HH_INTRA, D3_HH_INTRA, 1, 2) V(1, K3_HH_INTRA, D3_HH_INTRA, 4, 5) V(2, K3_HH_INTRA, D3_HH_INTRA, 7, 8) \
    V(3, K3_OH_INTRA, D3_OH_INTRA, 0, 1) V(4, K3_OH_INTRA, D3_OH_INTRA, 0, 2) V(5, K3_OH_INTRA, D3_OH_INTRA, 3, 4) \
    V(6, K3_OH_INTRA, D3_OH_INTRA, 3, 5) V(7, K3_OH_INTRA, D3_OH_INTRA, 6, 7) V(8, K3_OH_INTRA, D3_OH_INTRA, 6, 8) \
    V(9, K3_HH, D3_HH, 1, 4)  V(10, K3_HH, D3_HH, 1, 5) V(11, K3_HH, D3_HH, 1, 7)
#define MBPOL_3B_VARS_B(V)                                                                                        \
    V(12, K3_HH, D3_HH, 1, 8)                                                                                       \
    V(13, K3_HH, D3_HH, 2, 4) V(14, K3_HH, D3_HH, 2, 5) V(15, K3_HH, D3_HH, 2, 7) V(16, K3_HH, D3_HH, 2, 8)         \
    V(17, K3_HH, D3_HH, 4, 7) V(18, K3_HH, D3_HH, 4, 8) V(19, K3_HH, D3_HH, 5, 7) V(20, K3_HH, D3_HH, 5, 8)         \
    V(21, K3_OH, D3_OH, 0, 4) V(22, K3_OH, D3_OH, 0, 5) V(23, K3_OH, D3_OH, 0, 7)
#define MBPOL_3B_VARS_C(V)                                                                                        \
    V(24, K3_OH, D3_OH, 0, 8)                                                                                       \
    V(25, K3_OH, D3_OH, 3, 1) V(26, K3_OH, D3_OH, 3, 2) V(27, K3_OH, D3_OH, 3, 7) V(28, K3_OH, D3_OH, 3, 8)         \
    V(29, K3_OH, D3_OH, 6, 1) V(30, K3_OH, D3_OH, 6, 2) V(31, K3_OH, D3_OH, 6, 4) V(32, K3_OH, D3_OH, 6, 5)         \
    V(33, K3_OO, D3_OO, 0, 3) V(34, K3_OO, D3_OO, 0, 6) V(35, K3_OO, D3_OO, 3, 6)
#define MBPOL_3B_VARS(V) MBPOL_3B_VARS_A(V) MBPOL_3B_VARS_B(V) MBPOL_3B_VARS_C(V)

// Gather + image one trimer (nm).  Ob and Oc are both imaged relative to Oa, exactly like imageMolecules
// (MBPolReferenceTwoBodyForce.cpp:92,101) -- rbc is therefore not necessarily the minimum image; reproduced on purpose.
// Returns false when the reference early-outs (any O-O < 2 A, MBPolReferenceThreeBodyForce.cpp:165-166).
__device__ __forceinline__ bool trimer_sites(const double* __restrict__ pos, const int32_t* __restrict__ idx, int ma,
                                             int mb, int mc, const BoxArg& box, V3 (&s)[9], V3& rab, V3& rac, V3& rbc,
                                             double& dab, double& dac, double& dbc) {
    load_water(pos, idx, ma, 1.0, s[0], s[1], s[2]);
    load_water(pos, idx, mb, 1.0, s[3], s[4], s[5]);
    load_water(pos, idx, mc, 1.0, s[6], s[7], s[8]);
    if (box.periodic) {
        const double bx = box.len[0], by = box.len[1], bz = box.len[2];
        image_to(bx, by, bz, s[0], s[1]);
        image_to(bx, by, bz, s[0], s[2]);
        image_to(bx, by, bz, s[0], s[3]);
        image_to(bx, by, bz, s[3], s[4]);
        image_to(bx, by, bz, s[3], s[5]);
        image_to(bx, by, bz, s[0], s[6]);
        image_to(bx, by, bz, s[6], s[7]);
        image_to(bx, by, bz, s[6], s[8]);
    }
    rab = (s[0] - s[3]) * NM_TO_A;
    rac = (s[0] - s[6]) * NM_TO_A;
    rbc = (s[3] - s[6]) * NM_TO_A;
    dab = sqrt(dot(rab, rab));
    dac = sqrt(dot(rac, rac));
    dbc = sqrt(dot(rbc, rbc));
    return !(dab < 2.0 || dac < 2.0 || dbc < 2.0);
}

__global__ void __launch_bounds__(128) trimer_vars_kernel(const double* __restrict__ pos,
                                                          const int32_t* __restrict__ idx,
                                                          const int32_t* __restrict__ trips, int64_t first, int64_t n,
                                                          int64_t ld, BoxArg box, double* __restrict__ x) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int32_t* tr = trips + 3 * (first + t);
    V3 s[9], rab, rac, rbc;
    double dab, dac, dbc;
    const bool active = trimer_sites(pos, idx, tr[0], tr[1], tr[2], box, s, rab, rac, rbc, dab, dac, dbc);
#define V(i, k, r0, p, q)                                                \
    {                                                                    \
        double xi = 0.0;                                                 \
        if (active) {                                                    \
            const V3 d = (s[p] - s[q]) * NM_TO_A;                        \
            const double d2 = dot(d, d);                                 \
            const double ri = rsqrt(d2); /* see dimer_vars_kernel */     \
            const double r = d2 * ri;                                    \
            xi = exp(-(k) * (r - (r0)));                                 \
        }                                                                \
        x[(size_t)(i) * ld + t] = xi;                                    \
    }
    MBPOL_3B_VARS(V)
#undef V
}

__global__ void __launch_bounds__(128, 3) trimer_forces_kernel(const double* __restrict__ pos,
                                                               const int32_t* __restrict__ idx,
                                                               const int32_t* __restrict__ trips, int64_t first,
                                                               int64_t n, int64_t ld, BoxArg box,
                                                               const double* __restrict__ poly /* E, x_i dE/dx_i */,
                                                               double* __restrict__ e_out, double* __restrict__ f_out) {
    extern __shared__ double dyn_smem[];
    double* sh = dyn_smem;                              // [128][27]: 9 sites per thread, then its 9 x 3 force block
    double* col = dyn_smem + 128 * 27 + threadIdx.x;    // [1 + NV3][128]: this thread's polynomial outputs
    const int64_t cta_first = blockIdx.x * (int64_t)blockDim.x;
    const int64_t t = cta_first + threadIdx.x;
    if (t < n) stage_poly<NV3>(poly, ld, t, col);
    SiteRow<9> row{sh + 27 * threadIdx.x};
    V3 F[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) F[c] = {0.0, 0.0, 0.0};
    V3 rab, rac, rbc;
    double dab = 0.0, dac = 0.0, dbc = 0.0, e = 0.0;
    bool active = false;
    if (t < n) {
        const int32_t* tr = trips + 3 * (first + t);
        V3 s[9];
        active = trimer_sites(pos, idx, tr[0], tr[1], tr[2], box, s, rab, rac, rbc, dab, dac, dbc);
#pragma unroll
        for (int c = 0; c < 9; ++c) row.put(c, s[c]);
    }
    __pipeline_wait_prior(0);
    if (active) {
        // switching functions first (MBPolReferenceThreeBodyForce.cpp:106-120, 265-277): their gradient seeds the three
        // oxygens, and only `unit` stays live across the variable batches
        double gab, gac, gbc;
        const double sab = f_switch(dab, R3I, R3F, gab);
        const double sac = f_switch(dac, R3I, R3F, gac);
        const double sbc = f_switch(dbc, R3I, R3F, gbc);
        const double sw = sab * sac + sab * sbc + sac * sbc;
        const double E_poly = col[0];
        gab *= (sac + sbc) * E_poly / dab;
        gac *= (sab + sbc) * E_poly / dac;
        gbc *= (sab + sac) * E_poly / dbc;
        const double c = CAL2JOULE * -NM_TO_A;
        F[0] = (rab * gab + rac * gac) * c;
        F[3] = (rbc * gbc - rab * gab) * c;
        F[6] = (rac * gac + rbc * gbc) * -c;
        e = sw * E_poly * CAL2JOULE;
        // g_var (MBPolReferenceThreeBodyForce.cpp:81-104): gg = -k g x / d, then * cal2joule * 10 * -1; d in Angstrom
        const double unit = sw * (CAL2JOULE * 10.0 * -1.0) * NM_TO_A;
#define V(i, k, r0, p, q)                                                              \
    {                                                                                  \
        const V3 d = row.get(p) - row.get(q);                   /* nm */                \
        const double ri = rsqrt(dot(d, d)) * (1.0 / NM_TO_A);   /* 1 / Angstrom */      \
        const double xg = col[((i) + 1) * 128];                                        \
        const V3 dg = d * (-(k) * xg * ri * unit);                                     \
        F[p] += dg;                                                                    \
        F[q] -= dg;                                                                    \
    }
        MBPOL_3B_VARS(V)
#undef V
    }
    if (t < n) e_out[t] = e;
    // a thread's force block reuses its own site row (same 27 doubles): no barrier needed before the overwrite
#pragma unroll
    for (int a = 0; a < 9; ++a) row.put(a, F[a]);
    store_blocks<27, 128>(f_out, cta_first, n, sh);      // single barrier call site for the whole CTA
}

inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }

}  // namespace

void launch_dimer_vars(const double* pos, const int32_t* idx, const int32_t* pairs, int64_t first, int64_t n, int64_t ld,
                       BoxArg box, double* x, cudaStream_t s) {
    if (n > 0) dimer_vars_kernel<<<grid_for(n, 128), 128, 0, s>>>(pos, idx, pairs, first, n, ld, box, x); MBPOL_COUNT_LAUNCH(1);
}
void launch_trimer_vars(const double* pos, const int32_t* idx, const int32_t* trips, int64_t first, int64_t n, int64_t ld,
                        BoxArg box, double* x, cudaStream_t s) {
    if (n > 0) trimer_vars_kernel<<<grid_for(n, 128), 128, 0, s>>>(pos, idx, trips, first, n, ld, box, x); MBPOL_COUNT_LAUNCH(1);
}
void launch_dimer_forces(const double* pos, const int32_t* idx, const int32_t* pairs, int64_t first, int64_t n, int64_t ld,
                         BoxArg box, const double* poly, double* e, double* f, cudaStream_t s) {
    constexpr int smem = sizeof(double) * 128 * (30 + 1 + NV2);
    static const cudaError_t once = cudaFuncSetAttribute(dimer_forces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    (void)once;
    if (n > 0)
        dimer_forces_kernel<<<grid_for(n, 128), 128, smem, s>>>(pos, idx, pairs, first, n, ld, box, poly, e, f); MBPOL_COUNT_LAUNCH(1);
}
void launch_trimer_forces(const double* pos, const int32_t* idx, const int32_t* trips, int64_t first, int64_t n,
                          int64_t ld, BoxArg box, const double* poly, double* e, double* f, cudaStream_t s) {
    constexpr int smem = sizeof(double) * 128 * (27 + 1 + NV3);
    static const cudaError_t once = cudaFuncSetAttribute(trimer_forces_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    (void)once;
    if (n > 0)
        trimer_forces_kernel<<<grid_for(n, 128), 128, smem, s>>>(pos, idx, trips, first, n, ld, box, poly, e, f); MBPOL_COUNT_LAUNCH(1);
}

}  // namespace mbpol
