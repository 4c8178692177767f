/*
 * kde_oracle.c -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * Plain-C, fp64, single-purpose CPU restatement of the KDE discrepancy that
 * PopPUNK-mod evaluates for every simulated parameter set.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product (poppunk-mod_b200/) never links, imports or calls it.
 *
 * What is restated, and from where (paths relative to /root/reference unless
 * prefixed sklearn: / scipy: -- those are the third-party wheels the reference
 * calls; pinned scikit-learn 1.7.0 / scipy 1.11.4 / numpy 1.26.4 in
 * environment.yml:206,207,167; validated here against 1.9.0 / 1.18.1 / 2.3.5):
 *
 *   oracle_fit_scaler   KDE_distance.py:81-82  MinMaxScaler().fit(vstack([df1,df2]))
 *                       sklearn:preprocessing/_data.py:476-545 (nanmin/nanmax,
 *                       scale_=1/handle_zeros(range), min_=-data_min*scale_),
 *                       :101-133 (_handle_zeros_in_scale: range<10*eps -> 1)
 *   oracle_transform    KDE_distance.py:85-86, sklearn:_data.py:574-575 (X*=scale_; X+=min_)
 *                       then KDE_distance.py:53 (X[isnan(X)] = 0.0, AFTER scaling)
 *   oracle_linspace     KDE_distance.py:17-23 get_grid: xy[k]=(lin[k/R], lin[k%R])
 *   oracle_kde_dense /  KDE_distance.py:48-64 -> sklearn:neighbors/_kde.py:252-288 ->
 *   oracle_kde_window   sklearn:neighbors/_binary_tree.pxi.tp:389-394 (epanechnikov,
 *                       strict dist<h, dist = sqrt(rdist)), :448-476 (log kernel norm),
 *                       _kde.py:287 (- log N).  With rtol=atol=0 the tree visits every
 *                       in-support point, i.e. the result is this exact sum (plus the
 *                       tree's ~1e-12 log-space residue).
 *   oracle_pow_norm     KDE_distance.py:71-76  z**gamma ; z+=eps ; z/=z.sum()
 *   oracle_js           KDE_distance.py:129 -> scipy:spatial/distance.py:1278-1290
 *   oracle_discrepancy  KDE_distance.py:104-130 KDE_JS_divergence (log option :105-106)
 *
 * Parity pinning: the reference ships no tests or golden values for this path
 * (SURVEY.md section 4).  The oracle is pinned against outputs of the unmodified
 * reference module imported in the build container (oracle/make_golden.py ->
 * tests/golden/ (npz files); tests/test_oracle.py).
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off (oracle/Makefile).
 * -ffast-math / FMA contraction must NOT be used: separately rounded steps and
 * strict IEEE semantics are part of the contract.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define KERNEL_EPANECHNIKOV 0
#define KERNEL_GAUSSIAN 1

/* Neumaier compensated accumulator: the oracle should be closer to the exact
 * real-number sum than either implementation under test. */
typedef struct { double s, c; } ksum_t;
static inline void ksum_add(ksum_t *k, double x) {
    double t = k->s + x;
    if (fabs(k->s) >= fabs(x)) k->c += (k->s - t) + x; else k->c += (x - t) + k->s;
    k->s = t;
}
static inline double ksum_get(const ksum_t *k) { return k->s + k->c; }

/* optional df := log(df + eps), KDE_distance.py:105-106 */
static inline double pre(double v, int log_flag, double eps) {
    return log_flag ? log(v + eps) : v;
}

/* Joint NaN-ignoring min/max over both sets, then sklearn's scale_/min_.
 * Returns 0, or 1 if a non-NaN infinity is present (sklearn raises ValueError). */
int oracle_fit_scaler(const double *a, int64_t na, const double *b, int64_t nb,
                      int log_flag, double eps, double scale[2], double min_[2]) {
    int has_inf = 0;
    for (int c = 0; c < 2; ++c) {
        double mn = NAN, mx = NAN;
        for (int s = 0; s < 2; ++s) {
            const double *p = s ? b : a;
            int64_t n = s ? nb : na;
            for (int64_t i = 0; i < n; ++i) {
                double v = pre(p[2 * i + c], log_flag, eps);
                if (isnan(v)) continue;
                if (isinf(v)) has_inf = 1;
                if (isnan(mn) || v < mn) mn = v;
                if (isnan(mx) || v > mx) mx = v;
            }
        }
        double range = mx - mn;
        /* _handle_zeros_in_scale: scale[scale < 10*eps] = 1.0 (NaN compares false) */
        if (range < 10.0 * DBL_EPSILON) range = 1.0;
        scale[c] = 1.0 / range;
        min_[c] = 0.0 - mn * scale[c];
    }
    return has_inf;
}

void oracle_transform(const double *x, int64_t n, int log_flag, double eps,
                      const double scale[2], const double min_[2], double *out) {
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < 2; ++c) {
            double v = pre(x[2 * i + c], log_flag, eps);
            v = v * scale[c];           /* two separately rounded steps, as numpy does */
            v = v + min_[c];
            out[2 * i + c] = isnan(v) ? 0.0 : v;
        }
}

/* numpy.linspace(gmin, gmax, R): arange(R)*step + start, last point forced to stop */
void oracle_linspace(double gmin, double gmax, int R, double *lin) {
    if (R == 1) { lin[0] = gmin; return; }
    double step = (gmax - gmin) / (double)(R - 1);
    for (int i = 0; i < R; ++i) lin[i] = (double)i * step + gmin;
    lin[R - 1] = gmax;
}

static double kernel_norm(double h, int kernel) {
    /* _log_kernel_norm for d = 2: -(factor) - 2 log h */
    const double LOG_PI = 1.1447298858494001741434273513531;
    const double LOG_2PI = 1.8378770664093454835606594728112;
    double factor;
    if (kernel == KERNEL_GAUSSIAN) factor = 0.5 * 2 * LOG_2PI;
    else factor = (0.5 * 2 * LOG_PI - lgamma(0.5 * 2 + 1)) + log(2.0 / (2 + 2.0));
    return exp(-factor - 2.0 * log(h));
}

static inline double kernel_term(double dx, double dy, double h, int kernel) {
    double dist = sqrt(dx * dx + dy * dy);     /* euclidean dist(): sqrt of rdist */
    if (kernel == KERNEL_GAUSSIAN) return exp(-0.5 * (dist * dist) / (h * h));
    if (dist < h) return 1.0 - (dist * dist) / (h * h);
    return 0.0;
}

/* Definitional brute force: every node against every point.  dens is (R*R),
 * node k = ix*R + iy sits at (lin[ix], lin[iy]) with column 0 on the slow axis. */
void oracle_kde_dense(const double *X, int64_t n, double gmin, double gmax, int R,
                      double h, int kernel, double *dens) {
    double *lin = (double *)malloc(sizeof(double) * (size_t)R);
    oracle_linspace(gmin, gmax, R, lin);
    double knorm = kernel_norm(h, kernel);
    int64_t G = (int64_t)R * R;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < G; ++k) {
        double gx = lin[k / R], gy = lin[k % R];
        ksum_t acc = {0.0, 0.0};
        for (int64_t i = 0; i < n; ++i)
            ksum_add(&acc, kernel_term(gx - X[2 * i], gy - X[2 * i + 1], h, kernel));
        dens[k] = ksum_get(&acc) * knorm / (double)n;
    }
    free(lin);
}

/* Same sum, visiting only the node window that can be within h of each point
 * (window widened by one node each side; the strict dist<h test still decides).
 * Epanechnikov only.  Row-parallel so each node's points are added in input order. */
void oracle_kde_window(const double *X, int64_t n, double gmin, double gmax, int R,
                       double h, double *dens) {
    double *lin = (double *)malloc(sizeof(double) * (size_t)R);
    oracle_linspace(gmin, gmax, R, lin);
    double knorm = kernel_norm(h, KERNEL_EPANECHNIKOV);
    double step = (R > 1) ? (gmax - gmin) / (double)(R - 1) : 1.0;
    int64_t G = (int64_t)R * R;
    ksum_t *acc = (ksum_t *)calloc((size_t)G, sizeof(ksum_t));
#pragma omp parallel for schedule(static)
    for (int ix = 0; ix < R; ++ix) {
        double gx = lin[ix];
        for (int64_t i = 0; i < n; ++i) {
            double dx = gx - X[2 * i];
            if (!(fabs(dx) < h)) continue;
            double y = X[2 * i + 1];
            int lo = (int)floor((y - h - gmin) / step) - 1;
            int hi = (int)ceil((y + h - gmin) / step) + 1;
            if (lo < 0) lo = 0;
            if (hi > R - 1) hi = R - 1;
            for (int iy = lo; iy <= hi; ++iy) {
                double t = kernel_term(dx, lin[iy] - y, h, KERNEL_EPANECHNIKOV);
                if (t != 0.0) ksum_add(&acc[(int64_t)ix * R + iy], t);
            }
        }
    }
    for (int64_t k = 0; k < G; ++k) dens[k] = ksum_get(&acc[k]) * knorm / (double)n;
    free(acc);
    free(lin);
}

/* z = dens**gamma ; z += eps ; z /= sum(z)   (KDE_distance.py:71-76) */
void oracle_pow_norm(const double *dens, int64_t G, double gamma, double eps, double *z) {
    ksum_t s = {0.0, 0.0};
    for (int64_t k = 0; k < G; ++k) {
        z[k] = pow(dens[k], gamma) + eps;
        ksum_add(&s, z[k]);
    }
    double tot = ksum_get(&s);
    for (int64_t k = 0; k < G; ++k) z[k] = z[k] / tot;
}

static inline double rel_entr(double x, double y) {
    if (isnan(x) || isnan(y)) return NAN;
    if (x > 0.0 && y > 0.0) return x * log(x / y);
    if (x == 0.0 && y >= 0.0) return 0.0;
    return INFINITY;
}

/* scipy.spatial.distance.jensenshannon(p, q), natural log, returns the DISTANCE */
double oracle_js(const double *p, const double *q, int64_t G) {
    ksum_t sp = {0, 0}, sq = {0, 0};
    for (int64_t k = 0; k < G; ++k) { ksum_add(&sp, p[k]); ksum_add(&sq, q[k]); }
    double tp = ksum_get(&sp), tq = ksum_get(&sq);
    ksum_t l = {0, 0}, r = {0, 0};
    for (int64_t k = 0; k < G; ++k) {
        double pk = p[k] / tp, qk = q[k] / tq;
        double m = (pk + qk) / 2.0;
        ksum_add(&l, rel_entr(pk, m));
        ksum_add(&r, rel_entr(qk, m));
    }
    double js = ksum_get(&l) + ksum_get(&r);
    return sqrt(js / 2.0);
}

/* KDE_JS_divergence end to end.  Optional outputs (nullable): z1,z2 = normalised
 * probability vectors; d1,d2 = raw densities exp(score_samples).  dense != 0 uses
 * the brute-force density (required for the gaussian kernel).
 * Returns 0 ok, 1 infinity in input (sklearn would raise), 2 bad args. */
int oracle_discrepancy(const double *a, int64_t na, const double *b, int64_t nb,
                       int R, double h, double gamma, double eps, int log_flag,
                       int kernel, int dense, double *out_js,
                       double *z1, double *z2, double *d1, double *d2) {
    if (na <= 0 || nb <= 0 || R < 1 || !(h > 0)) return 2;
    double scale[2], min_[2];
    if (oracle_fit_scaler(a, na, b, nb, log_flag, eps, scale, min_)) return 1;
    int64_t G = (int64_t)R * R;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(2 * (na > nb ? na : nb)));
    double *dens[2], *z[2];
    for (int s = 0; s < 2; ++s) {
        dens[s] = (double *)malloc(sizeof(double) * (size_t)G);
        z[s] = (double *)malloc(sizeof(double) * (size_t)G);
        const double *p = s ? b : a;
        int64_t n = s ? nb : na;
        oracle_transform(p, n, log_flag, eps, scale, min_, buf);
        if (dense || kernel != KERNEL_EPANECHNIKOV)
            oracle_kde_dense(buf, n, 0.0, 1.0, R, h, kernel, dens[s]);
        else
            oracle_kde_window(buf, n, 0.0, 1.0, R, h, dens[s]);
        oracle_pow_norm(dens[s], G, gamma, eps, z[s]);
    }
    *out_js = oracle_js(z[0], z[1], G);
    if (z1) memcpy(z1, z[0], sizeof(double) * (size_t)G);
    if (z2) memcpy(z2, z[1], sizeof(double) * (size_t)G);
    if (d1) memcpy(d1, dens[0], sizeof(double) * (size_t)G);
    if (d2) memcpy(d2, dens[1], sizeof(double) * (size_t)G);
    for (int s = 0; s < 2; ++s) { free(dens[s]); free(z[s]); }
    free(buf);
    return 0;
}
