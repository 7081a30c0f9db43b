/* Minimal C client of the C ABI (include/mbpol_b200.h): two waters, 2B energy and forces with host buffers.
 *   gcc examples/capi_demo.c -Iinclude -Lmbpol_openmm_plugin_b200 -lmbpol_b200 -Wl,-rpath,$PWD/mbpol_openmm_plugin_b200 -o capi_demo
 * The dimer is the one of platforms/reference/tests/TestReferenceMBPolTwoBodyForce.cpp:99-105 (expected 6.14207815 kcal/mol). */
#include <stdio.h>

#include "mbpol_b200.h"

int main(void) {
    const double pos_A[6][3] = {{-1.516074336, -0.2023167650, 1.454672917}, {-0.6218989773, -0.6009430735, 1.572437625},
                                {-2.017613812, -0.4190350349, 2.239642849}, {-1.763651687, -0.3816594649, -1.300353949},
                                {-1.903851736, -0.4935677617, -0.3457810126}, {-2.527904158, -0.7613550077, -1.733803676}};
    double pos[18], force[18] = {0}, e[2];
    for (int i = 0; i < 18; ++i) pos[i] = pos_A[i / 3][i % 3] * 0.1;
    const int32_t idx[6] = {0, 1, 2, 3, 4, 5};
    mbpol_config cfg = {MBPOL_ABI_VERSION, 2, 6, idx, 10.0, 10.0, MBPOL_CUTOFF_NONPERIODIC, MBPOL_CUTOFF_NONPERIODIC, -1, 0};
    mbpol_ctx* ctx = NULL;
    if (mbpol_create(&ctx, &cfg) != MBPOL_OK) {
        fprintf(stderr, "mbpol_create: %s\n", mbpol_last_error(NULL));
        return 2;
    }
    if (mbpol_eval_host(ctx, pos, force, e, MBPOL_DO_2B) != MBPOL_OK) {
        fprintf(stderr, "mbpol_eval_host: %s\n", mbpol_last_error(ctx));
        return 3;
    }
    printf("E_2B = %.8f kcal/mol (reference test: 6.14207815)\n", e[0] / 4.184);
    for (int a = 0; a < 6; ++a)
        printf("grad[%d] = % .8f % .8f % .8f kcal/mol/A\n", a, -force[3 * a] / 41.84, -force[3 * a + 1] / 41.84, -force[3 * a + 2] / 41.84);
    mbpol_destroy(ctx);
    return 0;
}
