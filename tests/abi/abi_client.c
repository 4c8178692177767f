/* A plain-C caller of the drop-in boundary (include/kdejs.h): no Python, no torch, no CUDA headers.
 * usage: abi_client <input.bin> <resolution> <gamma> <eps> <log>
 * input.bin: int64 na, int64 nb, na*2 doubles, nb*2 doubles.  Prints "status <rc>" and "js <value>". */
#include <stdio.h>
#include <stdlib.h>
#include "kdejs.h"

int main(int argc, char **argv) {
    if (argc < 6) { fprintf(stderr, "usage\n"); return 64; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 65; }
    int64_t n[2];
    if (fread(n, sizeof(int64_t), 2, f) != 2) return 66;
    double *a = (double *)malloc(sizeof(double) * 2 * (size_t)n[0]);
    double *b = (double *)malloc(sizeof(double) * 2 * (size_t)n[1]);
    if (fread(a, sizeof(double), 2 * (size_t)n[0], f) != 2 * (size_t)n[0]) return 67;
    if (fread(b, sizeof(double), 2 * (size_t)n[1], f) != 2 * (size_t)n[1]) return 67;
    fclose(f);
    double js = -1.0;
    int rc = kdejs_discrepancy(0, a, n[0], b, n[1], atoi(argv[2]), 0.03, atof(argv[3]), atof(argv[4]),
                               atoi(argv[5]), KDEJS_KERNEL_EPANECHNIKOV, KDEJS_ALGO_BINNED, &js, NULL, NULL, NULL, NULL);
    printf("status %d\n", rc);
    if (rc != KDEJS_OK) printf("error %s\n", kdejs_last_error());
    else printf("js %.17g\n", js);
    /* a second entry point through the same ABI: batch of the same pair twice */
    if (rc == KDEJS_OK) {
        const double *sims[2] = {b, b};
        int64_t ns[2] = {n[1], n[1]};
        double out[2];
        rc = kdejs_discrepancy_batch(0, a, n[0], sims, ns, 2, atoi(argv[2]), 0.03, atof(argv[3]), atof(argv[4]),
                                     atoi(argv[5]), KDEJS_KERNEL_EPANECHNIKOV, KDEJS_ALGO_BINNED, out);
        printf("batch %d %.17g %.17g\n", rc, out[0], out[1]);
    }
    kdejs_shutdown();
    free(a); free(b);
    return rc == KDEJS_OK ? 0 : 1;
}
