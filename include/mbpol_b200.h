/* mbpol_b200.h -- C ABI of the B200-native MB-pol 2B+3B short-range many-body path.
 *
 * This is the drop-in boundary: everything a platform kernel of the reference plugin does for
 *   CalcMBPolTwoBodyForceKernel   (openmmapi/include/openmm/mbpolKernels.h:134-168)
 *   CalcMBPolThreeBodyForceKernel (openmmapi/include/openmm/mbpolKernels.h:173-207)
 * i.e. what ReferenceCalcMBPol{Two,Three}BodyForceKernel::initialize/execute/copyParametersToContext do
 * (platforms/reference/src/MBPolReferenceKernels.cpp:341-410, 430-497) and what the to-be-replaced
 * CudaCalcMBPol{Two,Three}BodyForceKernel do (platforms/cuda/src/CudaMBPolKernels.cpp:174-304, 1722-2005),
 * is reachable through the plain-C entry points below.  No C++ / torch / OpenMM types cross this boundary: plain
 * pointers, sizes and a CUDA stream handle passed as void*.  Every function returns 0 on success or a negative
 * MBPOL_E* code; mbpol_last_error(ctx) gives the message (the C++ adapter turns it into OpenMM::OpenMMException).
 *
 * Units follow OpenMM: positions / box / cutoffs in nm, energies in kJ/mol, forces in kJ/mol/nm.
 * One context per OpenMM Context; a context is not re-entrant.
 */
#ifndef MBPOL_B200_H
#define MBPOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBPOL_ABI_VERSION 2

/* error codes */
#define MBPOL_OK            0
#define MBPOL_EINVAL       (-1)   /* bad argument */
#define MBPOL_ECUDA        (-2)   /* CUDA runtime / driver error */
#define MBPOL_EBOX         (-3)   /* periodic box smaller than twice the cutoff (MBPolReferenceKernels.cpp:384-391) */
#define MBPOL_ECHANGED     (-4)   /* number of molecules changed (MBPolReferenceKernels.cpp:402-403) */
#define MBPOL_ENOMEM       (-5)
#define MBPOL_ECOMM        (-6)   /* NCCL missing or a collective failed */

/* nonbonded method: values of MBPolTwoBodyForce::NonbondedMethod (openmmapi/include/openmm/MBPolTwoBodyForce.h:62-74) */
#define MBPOL_NOCUTOFF            0   /* treated as CutoffNonPeriodic with an infinite cutoff */
#define MBPOL_CUTOFF_PERIODIC     1
#define MBPOL_CUTOFF_NONPERIODIC  2

/* mbpol_eval flags */
#define MBPOL_DO_2B        1u     /* evaluate the MBPolTwoBodyForce term   */
#define MBPOL_DO_3B        2u     /* evaluate the MBPolThreeBodyForce term */
#define MBPOL_KEEP_LISTS   4u     /* reuse the tuple lists of the previous call (default: rebuild, like the reference) */
#define MBPOL_OVERWRITE_FORCES 8u /* mbpol_eval_host only: `force` is output-only (not read, not added to): saves the
                                     host->device copy of the force array */

typedef struct mbpol_ctx mbpol_ctx;

/* What MBPol{Two,Three}BodyForce carry (MBPolTwoBodyForce.h:57-166): per-molecule particle indices {O,H1,H2},
 * a cutoff and a nonbonded method each.  The PIP coefficients and nonlinear parameters are baked into the library
 * (they are compile-time constants in the reference too, mbpol_2body_constants.h:33). */
typedef struct mbpol_config {
    int32_t        abi_version;      /* MBPOL_ABI_VERSION */
    int32_t        n_molecules;      /* number of addParticle() calls */
    int32_t        n_particles;      /* length of the position / force arrays (>= 3*n_molecules; M sites are skipped) */
    const int32_t* particle_indices; /* n_molecules x 3 global indices {O,H1,H2}, copied */
    double         cutoff_2b_nm;     /* mbpol.xml: 0.65 */
    double         cutoff_3b_nm;     /* mbpol.xml: 0.45 */
    int32_t        method_2b;        /* MBPOL_NOCUTOFF | MBPOL_CUTOFF_PERIODIC | MBPOL_CUTOFF_NONPERIODIC */
    int32_t        method_3b;
    int32_t        device;           /* CUDA device ordinal, -1 = current device */
    int32_t        reserved;
} mbpol_config;

/* counters of the last mbpol_eval call */
typedef struct mbpol_stats {
    int64_t n_pairs;        /* listed dimers  (O-O <= cutoff_2b); -1 when the lists of this evaluation were SLICED (multi-  */
    int64_t n_triplets;     /* GPU: only this rank's tuples exist, the box-wide totals are unknown); trimers: chain set      */
    int64_t n_pairs_owned;  /* dimers / trimers evaluated by THIS context (owner range)  */
    int64_t n_triplets_owned;
    float   ms_lists, ms_2b, ms_3b, ms_reduce;   /* device time of the stages (CUDA events), 0 if timing disabled; when the two
                                                    forces disagree on periodicity (two list builds) they describe the 3B pass */
    float   ms_pip_2b, ms_pip_3b;                /* the polynomial kernels alone (the roofline kernels)             */
    int64_t n_launches;     /* kernels launched by the last mbpol_eval (library-internal CUB passes counted once each) */
    float   ms_bcast, ms_comm_reduce;            /* mbpol_eval_distributed: the two collectives (CUDA events)        */
    int32_t lists_sliced;   /* 1: the lists of this evaluation hold only this rank's tuples (n_pairs / n_triplets = -1) */
    int32_t lists_reused;   /* 1: the neighbour rows of an earlier evaluation were reused (Verlet skin / KEEP_LISTS)     */
    int64_t n_repartitions; /* full list builds with a fresh owner balance since mbpol_set_partition                  */
    int64_t n_row_builds;   /* neighbour-row builds since the context was created (skin reuse skips them)             */
} mbpol_stats;

/* replaces CalcMBPol*ForceKernel::initialize (MBPolReferenceKernels.cpp:341-360, 430-449) */
int  mbpol_create(mbpol_ctx** out, const mbpol_config* cfg);
void mbpol_destroy(mbpol_ctx* ctx);
const char* mbpol_last_error(const mbpol_ctx* ctx);     /* ctx may be NULL: error of the last failed mbpol_create */

/* replaces CalcMBPol*ForceKernel::copyParametersToContext (MBPolReferenceKernels.cpp:401-415, 485-497):
 * new particle indices for the same number of molecules; MBPOL_ECHANGED if the count differs */
int  mbpol_update_particles(mbpol_ctx* ctx, int32_t n_molecules, const int32_t* particle_indices);

/* rectangular periodic box edge lengths (extractBoxSize, MBPolReferenceKernels.cpp:63-66) */
int  mbpol_set_box(mbpol_ctx* ctx, const double box_nm[3]);

/* Verlet skin.  0 (default) = neighbour rows are rebuilt from the cell grid on every evaluation, like
 * the reference (MBPolReferenceKernels.cpp:379-383).  skin_nm > 0: the cell scan builds rows out to (largest cutoff +
 * skin), keeps them while no oxygen has moved more than skin / 2 (minimum image) since, and every evaluation filters
 * them with the exact d^2 <= cutoff^2 test at the current positions -- the pair / triplet lists are the reference's sets
 * at every step either way.  Applies to unpartitioned contexts (world = 1); mbpol_stats.lists_reused / n_row_builds tell
 * what happened. */
int  mbpol_set_skin(mbpol_ctx* ctx, double skin_nm);

/* multi-GPU owner-molecule partition: this context evaluates only the tuples whose owner (highest molecule index I)
 * falls in slice `rank` of `world` slices balanced by 30,601*dimers + 111,818*trimers per owner.
 * world = 1 (default) evaluates everything.  Forces / energies returned are then partial sums to be all-reduced.
 * The first evaluation after this call builds the full lists and fixes the rank's owner range; later evaluations
 * rebuild, every step, only the neighbour rows that range needs (its owners + their 0.45 nm halo), and
 * mbpol_get_pairs / mbpol_get_triplets then return the rank's own tuples.  Every MBPOL_REPARTITION_EVERY (env,
 * default 50; 0 = always build everything) evaluations the full lists and a fresh balance are computed again.  All
 * ranks must therefore evaluate in lockstep on the same positions (the broadcast of SURVEY.md 8e does that). */
int  mbpol_set_partition(mbpol_ctx* ctx, int32_t rank, int32_t world);

/* replaces CalcMBPol{Two,Three}BodyForceKernel::execute (MBPolReferenceKernels.cpp:362-399, 451-483).
 *   pos_nm : DEVICE pointer, n_particles x 3 doubles
 *   force  : DEVICE pointer, n_particles x 3 doubles, the path ADDS its forces (like the reference); may be NULL
 *   energy : HOST pointer to 2 doubles {E_2B, E_3B} in kJ/mol; may be NULL (no device->host sync is then made)
 *   stream : cudaStream_t / CUstream as void*; NULL = legacy default stream */
int  mbpol_eval(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, uint32_t flags, void* stream);

/* same call with HOST buffers: copies positions in, forces out (force is ADDED into), synchronous.
 * This is the entry the Reference-platform-style adapter and the end-to-end benchmark use. */
int  mbpol_eval_host(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, uint32_t flags);

/* replaces CalcMBPolOneBodyForceKernel::execute (openmmapi/include/openmm/mbpolKernels.h:47-82,
 * platforms/reference/src/MBPolReferenceKernels.cpp:101-114; MBPolReferenceOneBodyForce.cpp:70-210): the intramolecular
 * Partridge-Schwenke surface of every molecule of the context (the MBPolOneBodyForce::addOneBody list is the context's
 * particle_indices).  periodic = MBPolOneBodyForce::Periodic (hydrogens imaged to their own oxygen with the box of
 * mbpol_set_box).  pos_nm / force: DEVICE pointers as in mbpol_eval (forces are ADDED; every particle must belong to
 * one molecule only); energy: HOST pointer to 1 double (kJ/mol) or NULL.  With a partition (mbpol_set_partition) rank r
 * evaluates molecules [n*r/world, n*(r+1)/world). */
int  mbpol_eval_onebody(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, int32_t periodic,
                        void* stream);
/* same with HOST buffers (synchronous) */
int  mbpol_eval_onebody_host(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, int32_t periodic);

/* OpenMM CUDA force-buffer convention (platforms/cuda/src/kernels/twobodyForce.cu:442-457): int64 fixed point,
 * value * 2^32, SoA [x: padded][y: padded][z: padded].  Adds the forces of the last mbpol_eval (which must have been
 * called with force == NULL or not) into omm_force for the particles of this context. DEVICE pointer. */
int  mbpol_add_forces_fixed_point(mbpol_ctx* ctx, long long* omm_force, int32_t padded_num_atoms, void* stream);

/* tuple lists of the last evaluation in canonical order (ascending I, then J[, then K]; I > J > K), for bit-exact
 * comparison with computeNeighborListVoxelHash / computeThreeNeighborListVoxelHash
 * (platforms/reference/src/ReferenceThreeNeighborList.cpp:181-226).  HOST pointers.  Call with out == NULL to get *n. */
int  mbpol_get_pairs(mbpol_ctx* ctx, int32_t* out /* n x 2 */, size_t* n);
int  mbpol_get_triplets(mbpol_ctx* ctx, int32_t* out /* n x 3 */, size_t* n);

int  mbpol_get_stats(const mbpol_ctx* ctx, mbpol_stats* out);
int  mbpol_set_timing(mbpol_ctx* ctx, int32_t enabled);   /* record per-stage CUDA events (adds a sync per eval) */

/* ---- energies without a host synchronisation -------------------------------------------------------------------
 * mbpol_eval with energy == NULL leaves {E_2B, E_3B} of the call in device memory.  mbpol_energy_device returns that
 * DEVICE pointer (4 doubles: E_2B, E_3B, E_1B, E_dispersion of the last evaluations; valid in stream order after the
 * call);
 * mbpol_get_energy copies them to the host (synchronises `stream`). */
const double* mbpol_energy_device(const mbpol_ctx* ctx);
int  mbpol_get_energy(mbpol_ctx* ctx, double energy[3], void* stream);

/* ---- OpenMM CUDA platform buffers (the ownership row of SURVEY.md 8b) ---------------------------------------------
 * What CudaCalcMBPol{Two,Three}BodyForceKernel::execute works on (platforms/cuda/src/CudaMBPolKernels.cpp:278-304,
 * 1976-2005): cu.getPosq() (real4, or double4 with CudaPrecision=double), in OpenMM's REORDERED atom order
 * (cu.getAtomIndex(): slot -> original particle index), cu.getForce() (int64 fixed point, value * 2^32, SoA
 * [x: padded][y: padded][z: padded], same slot order) and cu.getEnergyBuffer() (element 0 receives += energy; double
 * in the double and mixed precision modes, float in single).  The particle indices given to mbpol_create are ORIGINAL
 * indices, as in the MBPol*Force objects.  All pointers DEVICE; atom_index may be NULL (identity order); force_buffer
 * and energy_buffer may be NULL.  Nothing is copied to the host and the stream is not synchronised: execute() returns
 * 0.0 on the CUDA platform (CudaMBPolKernels.cpp:303). */
int  mbpol_eval_openmm(mbpol_ctx* ctx, const void* posq, int32_t posq_is_double, const int32_t* atom_index,
                       int32_t num_atoms, long long* force_buffer, int32_t padded_num_atoms, void* energy_buffer,
                       int32_t energy_is_double, uint32_t flags, void* stream);

/* ---- multi-GPU behind the C ABI: NCCL broadcast / reduce (north_star (3)) ----------------------------
 * One process or thread per GPU, one context per rank.  Rank 0 obtains a 128-byte id (ncclGetUniqueId), the host
 * program ships it to the other ranks (MPI, a file, torch.distributed's store ...), every rank calls
 * mbpol_comm_init_rank (ncclCommInitRank on the context's device + mbpol_set_partition(rank, world)).
 * mbpol_comm_adopt uses an ncclComm_t the host program already owns instead (rank / world are read from it; it is not
 * destroyed with the context).  NCCL is bound at run time (libnccl.so.2); MBPOL_ECOMM if it is not there. */
#define MBPOL_COMM_ID_BYTES 128
int  mbpol_comm_unique_id(void* id_out /* MBPOL_COMM_ID_BYTES */);
int  mbpol_comm_init_rank(mbpol_ctx* ctx, const void* id, int32_t rank, int32_t world);
int  mbpol_comm_adopt(mbpol_ctx* ctx, void* nccl_comm);
int  mbpol_comm_destroy(mbpol_ctx* ctx);

/* One distributed evaluation, all ranks in lockstep on `stream`:
 *   ncclBroadcast(pos_nm) from rank 0  ->  lists + 2B + 3B of this rank's owner range  ->  ONE grouped NCCL reduce of
 *   the force array and the two energies to rank 0 (MBPOL_DIST_ALLREDUCE: all-reduce, every rank receives them, e.g.
 *   for a replicated integrator; MBPOL_DIST_NO_BCAST: positions are already identical on every rank).
 *   pos_nm : DEVICE, n_particles x 3; input on rank 0, overwritten by the broadcast elsewhere
 *   force  : DEVICE, n_particles x 3; OVERWRITTEN with the summed forces on the receiving rank(s); may be NULL elsewhere
 *   energy : HOST, 2 doubles {E_2B, E_3B}, written on the receiving rank(s); NULL = no host synchronisation */
#define MBPOL_DIST_ALLREDUCE 16u
#define MBPOL_DIST_NO_BCAST  32u
int  mbpol_eval_distributed(mbpol_ctx* ctx, double* pos_nm, double* force, double* energy, uint32_t flags,
                            void* stream);

/* The same with HOST buffers on rank 0 (the Reference-platform calling convention, extractPositions / extractForces,
 * MBPolReferenceKernels.cpp:51-61, for a multi-GPU run): pos_host (rank 0: n_particles x 3, ideally pinned) goes to the
 * device and is broadcast; the forces are reduced to rank 0 in four slices and every slice starts its way back to
 * force_host (rank 0, OVERWRITTEN) on a copy stream as soon as its reduce is done, so the device-to-host transfer hides
 * behind the remaining reduces.  Other ranks pass NULL for both.  Returns after the results are in host memory. */
int  mbpol_eval_distributed_host(mbpol_ctx* ctx, const double* pos_host, double* force_host, double* energy,
                                 uint32_t flags);

/* ---- TT-damped C6 dispersion ------------------------------------------------------------------
 * What python/mbpol.xml:37-108 builds as an OpenMM CustomNonbondedForce: for every pair of O / H atoms of DIFFERENT
 * molecules (M sites excluded) with minimum-image distance r <= cutoff_nm:  E = -C6 tt6(r) / r^6,
 * tt6 = 1 - exp(-d6 r) sum_{k=0..6} (d6 r)^k / k!, C6 / d6 per atom-class pair (mbpol.xml:53-82); no switching function,
 * no long-range correction.  The molecules are the context's particle_indices; pos_nm / force DEVICE pointers as in
 * mbpol_eval (forces are ADDED), energy HOST pointer to 1 double (kJ/mol) or NULL; periodic uses the box of mbpol_set_box
 * (MBPOL_EBOX if cutoff_nm > box / 2, OpenMM's own rule); cutoff_nm >= 1e9 means no cutoff.  With a partition
 * (mbpol_set_partition) the rank evaluates the molecule pairs of its owner range. */
int  mbpol_eval_dispersion(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, double cutoff_nm,
                           int32_t periodic, void* stream);
int  mbpol_eval_dispersion_host(mbpol_ctx* ctx, const double* pos_nm, double* force, double* energy, double cutoff_nm,
                                int32_t periodic);

/* ---- on-device NVE integrator -------------------------------------------------------------------
 * n_steps of velocity Verlet around the forces of this context, entirely on the device:
 *     v += dt/2 F/m ;  x += dt v ;  M sites placed ;  F = F_1B + F_2B + F_3B (x) ;  v += dt/2 F/m
 * (dt 0.2 fs and the masses of python/example_ini/mbpol_bulk_nve.ini:11 / python/mbpol.xml:3-4 are the caller's choice).
 *   pos, vel   : DEVICE n_particles x 3, updated in place (nm, nm/ps)
 *   mass       : DEVICE n_particles (amu); 0 = massless (virtual site): neither kicked nor drifted
 *   cfg->msite_index : HOST n_molecules particle indices of the molecules' M sites (-1: none) or NULL; placed after every
 *                drift with the average3 weights cfg->msite_weights (python/mbpol.xml:14: 0.573293118, 0.213353441 x 2)
 *   energy_log : HOST (n_steps + 1) x 2 doubles {E_pot, E_kin} in kJ/mol (row 0 = start), copied once at the end; NULL =
 *                no synchronisation at all
 * No positions, forces or energies cross the host boundary during the run; the tuple lists are rebuilt (or, with
 * mbpol_set_skin, filtered) every step like the reference does. */
#define MBPOL_NVE_ONEBODY 64u     /* cfg->flags: include the one-body term (mbpol_eval_onebody) */
typedef struct mbpol_nve_config {
    double         dt_ps;
    uint32_t       flags;             /* MBPOL_DO_2B | MBPOL_DO_3B | MBPOL_NVE_ONEBODY */
    int32_t        onebody_periodic;  /* MBPolOneBodyForce::Periodic */
    const int32_t* msite_index;
    double         msite_weights[3];
} mbpol_nve_config;
int  mbpol_nve_run(mbpol_ctx* ctx, double* pos, double* vel, const double* mass, const mbpol_nve_config* cfg,
                   int32_t n_steps, double* energy_log, void* stream);

/* ---- lower-level entry points used by the unit tests and the roofline benchmark ------------------------------- */

/* evaluate the bare PIP (poly_2b_v6x_eval: poly-2b-v6x.cpp:3 / poly_3b_v2x::eval: poly-3b-v2x.cpp:3) on n tuples.
 *   order : 2 or 3;  x : DEVICE [n_vars][ld] (31 or 36 variables, tuple index fastest)
 *   out   : DEVICE [1+n_vars][ld]: row 0 = E_poly (kcal/mol), rows 1.. = dE/dx */
int  mbpol_pip_eval(int order, const double* x, double* out, int64_t n, int64_t ld, void* stream);

/* per-tuple energies (kJ/mol) and force blocks (kJ/mol/nm; 6x3 / 9x3 per tuple in slot order Oa,Ha1,Ha2,Ob,...)
 * for explicit tuple lists over stride-3 molecules (pos = n_mol x 3 x 3).  All pointers DEVICE. */
int  mbpol_eval_tuples(int order, const double* pos_nm, int32_t n_mol, const int32_t* tuples, int64_t n_tuples,
                       const double box_nm[3], int periodic, double* e_each, double* f_each, void* stream);

/* register-resident DFMA loop: measures the FP64 FMA-pipe peak of the device this library runs on.
 * Returns achieved TFLOP/s in *tflops. */
int  mbpol_measure_fp64_peak(double* tflops, double* ms);

/* constants of the work model */
#define MBPOL_FLOP_PER_DIMER   30601   /* 10,707 mul + 19,894 add, poly-2b-v6x.cpp */
#define MBPOL_FLOP_PER_TRIMER  111818  /* 22,294 mul + 89,524 add, poly-3b-v2x.cpp */
int  mbpol_pip_instruction_count(int order);   /* FP64 instructions our evaluator issues per tuple */

#ifdef __cplusplus
}
#endif
#endif /* MBPOL_B200_H */
