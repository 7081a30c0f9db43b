/* Two GPUs through the C ABI alone (include/mbpol_b200.h): one thread per GPU, NCCL inside the library.
 *   gcc examples/capi_demo_2gpu.c -Iinclude -I/usr/local/cuda/include -Lmbpol_openmm_plugin_b200 -lmbpol_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -lpthread -Wl,-rpath,$PWD/mbpol_openmm_plugin_b200 -o capi_demo_2gpu
 * The water trimer of platforms/reference/tests/TestReferenceMBPolThreeBodyForce.cpp:97-107 is evaluated (a) on GPU 0 alone
 * and (b) partitioned by owner molecule over GPUs 0 and 1 with mbpol_eval_distributed (NCCL broadcast of the positions,
 * one grouped NCCL reduce of forces + energies); both must agree.  Needs 2 GPUs; exits 77 with fewer. */
#include <cuda_runtime.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <string.h>

#include "mbpol_b200.h"

static const double pos_A[9][3] = {
    {-1.516074336, -0.2023167650, 1.454672917},  {-0.6218989773, -0.6009430735, 1.572437625}, {-2.017613812, -0.4190350349, 2.239642849},
    {-1.763651687, -0.3816594649, -1.300353949}, {-1.903851736, -0.4935677617, -0.3457810126}, {-2.527904158, -0.7613550077, -1.733803676},
    {-0.5588472140, 2.006699172, -0.1392786582}, {-0.9411558180, 1.541226676, 0.6163293071},  {-0.9858551734, 1.567124294, -0.8830970941}};
static const int32_t idx[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
static char nccl_id[MBPOL_COMM_ID_BYTES];
static double result_force[27], result_energy[2];
static int status[2];

static mbpol_ctx* make_ctx(int device) {
    mbpol_config cfg = {MBPOL_ABI_VERSION, 3, 9, idx, 10.0, 10.0, MBPOL_CUTOFF_NONPERIODIC, MBPOL_CUTOFF_NONPERIODIC, device, 0};
    mbpol_ctx* ctx = NULL;
    if (mbpol_create(&ctx, &cfg) != MBPOL_OK) fprintf(stderr, "mbpol_create(device %d): %s\n", device, mbpol_last_error(NULL));
    return ctx;
}

static void* rank_main(void* arg) {
    const int rank = (int)(size_t)arg;
    status[rank] = 1;
    cudaSetDevice(rank);
    mbpol_ctx* ctx = make_ctx(rank);
    if (!ctx) return NULL;
    double *pos = NULL, *force = NULL, host_pos[27];
    cudaMalloc((void**)&pos, sizeof host_pos);
    cudaMalloc((void**)&force, sizeof host_pos);
    if (rank == 0) {                                   /* only rank 0 knows the positions: the library broadcasts them */
        for (int i = 0; i < 27; ++i) host_pos[i] = pos_A[i / 3][i % 3] * 0.1;
        cudaMemcpy(pos, host_pos, sizeof host_pos, cudaMemcpyHostToDevice);
    } else cudaMemset(pos, 0, sizeof host_pos);
    if (mbpol_comm_init_rank(ctx, nccl_id, rank, 2) != MBPOL_OK) {
        fprintf(stderr, "rank %d: %s\n", rank, mbpol_last_error(ctx));
        return NULL;
    }
    double e[2] = {0, 0};
    if (mbpol_eval_distributed(ctx, pos, force, rank == 0 ? e : NULL, MBPOL_DO_2B | MBPOL_DO_3B, NULL) != MBPOL_OK) {
        fprintf(stderr, "rank %d: %s\n", rank, mbpol_last_error(ctx));
        return NULL;
    }
    cudaDeviceSynchronize();
    if (rank == 0) {
        cudaMemcpy(result_force, force, sizeof result_force, cudaMemcpyDeviceToHost);
        memcpy(result_energy, e, sizeof e);
    }
    mbpol_stats st;
    mbpol_get_stats(ctx, &st);
    printf("rank %d evaluated %lld dimers + %lld trimers\n", rank, (long long)st.n_pairs_owned, (long long)st.n_triplets_owned);
    cudaFree(pos);
    cudaFree(force);
    mbpol_destroy(ctx);
    status[rank] = 0;
    return NULL;
}

int main(void) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 2) {
        printf("needs 2 GPUs (found %d)\n", ndev);
        return 77;
    }
    /* (a) single GPU, host buffers */
    double pos[27], f1[27] = {0}, e1[2];
    for (int i = 0; i < 27; ++i) pos[i] = pos_A[i / 3][i % 3] * 0.1;
    mbpol_ctx* one = make_ctx(0);
    if (!one || mbpol_eval_host(one, pos, f1, e1, MBPOL_DO_2B | MBPOL_DO_3B) != MBPOL_OK) return 2;
    mbpol_destroy(one);
    /* (b) two ranks */
    if (mbpol_comm_unique_id(nccl_id) != MBPOL_OK) {
        fprintf(stderr, "mbpol_comm_unique_id: %s\n", mbpol_last_error(NULL));
        return 3;
    }
    pthread_t th[2];
    for (size_t r = 0; r < 2; ++r) pthread_create(&th[r], NULL, rank_main, (void*)r);
    for (int r = 0; r < 2; ++r) pthread_join(th[r], NULL);
    if (status[0] || status[1]) return 4;
    double df = 0.0;
    for (int i = 0; i < 27; ++i) df = fmax(df, fabs(result_force[i] - f1[i]));
    printf("E_2B = %.10f / %.10f   E_3B = %.10f / %.10f kJ/mol (1 GPU / 2 GPUs), max |dF| = %.3e kJ/mol/nm\n", e1[0],
           result_energy[0], e1[1], result_energy[1], df);
    const int ok = fabs(result_energy[0] - e1[0]) <= 1e-9 * fabs(e1[0]) && fabs(result_energy[1] - e1[1]) <= 1e-9 * fabs(e1[1]) && df < 1e-6;
    printf(ok ? "Done\n" : "MISMATCH\n");
    return ok ? 0 : 1;
}
