/* A plain C99 host driving the multi-GPU entry points of include/klineprofiles_b200.h the way a single-process caller
 * (the reference's Zig host, XSPEC, Julia) would: one table replica per visible GPU (at most 2 here), a local group, a batch
 * of kline5 sets through klp_group_eval_batch, compared bit for bit with klp_eval_batch on one GPU.
 * usage: group_client <table.klpt|.fits> <n_sets>      exit 0 = identical, 77 = no CUDA device */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "klineprofiles_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc__ = (call);                                                           \
        if (rc__ != KLP_OK) {                                                        \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, klp_last_error());  \
            return 1;                                                                \
        }                                                                            \
    } while (0)

static double lcg(unsigned long long* s) { /* deterministic parameters, no libm needed */
    *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (double)(*s >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s table n_sets\n", argv[0]); return 2; }
    const long B = atol(argv[2]);
    const int n = 400, P = klp_model_nparams(KLP_KLINE5);
    int ndev = 0;
    if (klp_device_count(&ndev) != KLP_OK || ndev < 1) { printf("no CUDA device: skipped\n"); return 77; }
    if (ndev > 2) ndev = 2;
    klp_table* tables[2] = {NULL, NULL};
    for (int d = 0; d < ndev; ++d) CHECK(klp_table_load(argv[1], d, &tables[d]));
    klp_group* g = NULL;
    CHECK(klp_group_create_local(tables, ndev, &g));
    int n_local = 0, nranks = 0;
    CHECK(klp_group_size(g, &n_local, &nranks));
    double* params = malloc(sizeof(double) * (size_t)B * P);
    double* energy = malloc(sizeof(double) * (n + 1));
    double* one = malloc(sizeof(double) * (size_t)B * n);
    double* all = malloc(sizeof(double) * (size_t)B * n);
    unsigned long long seed = 42;
    for (long s = 0; s < B; ++s) {
        double* p = params + s * P;
        p[0] = 0.998 * lcg(&seed); p[1] = 5 + 80 * lcg(&seed); p[2] = 6 + lcg(&seed); p[3] = 1 + 9 * lcg(&seed);
        p[4] = 20 + 380 * lcg(&seed); p[5] = 5 + 95 * lcg(&seed); p[6] = 5 * lcg(&seed);
        for (int k = 7; k < P; ++k) p[k] = 5 * lcg(&seed);
    }
    for (int i = 0; i <= n; ++i) energy[i] = 0.1 + 9.9 * i / n;
    CHECK(klp_eval_batch(tables[0], KLP_KLINE5, KLP_F32, B, params, energy, n, NULL, 0, one));
    CHECK(klp_group_eval_batch(g, KLP_KLINE5, KLP_F32, B, params, energy, n, NULL, 0, all));
    const int same = memcmp(one, all, sizeof(double) * (size_t)B * n) == 0;
    double sum = 0;
    for (long i = 0; i < B * n; ++i) sum += all[i];
    printf("devices %d (group: %d local of %d ranks), %ld sets, sum of spectra %.6f, group == single GPU: %s\n", ndev, n_local, nranks, B,
           sum, same ? "yes" : "NO");
    klp_group_destroy(g);
    for (int d = 0; d < ndev; ++d) klp_table_destroy(tables[d]);
    free(params); free(energy); free(one); free(all);
    return same ? 0 : 1;
}
