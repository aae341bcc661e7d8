/* TEST INFRASTRUCTURE — CPU oracle for the VQ / LBG codebook path.  NOT product code.
 *
 * Plain-C fp64 restatement of the reference's vector-quantisation encoder and LBG
 * (binary-split k-means) codebook builder (DemianMArin/HMM_Training).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from this file (oracle/_build/libvq_oracle.so).
 *
 * Reference lines restated (paths relative to the reference root):
 *   - VQ encode ............ HMM/hmm_training.py:95-118   (dims 1..12, strict '<')
 *   - euclidian_distance ... CodeVector/codevector_functions.py:82-87
 *   - new_epsilon_centroids  CodeVector/codevector_functions.py:383-411 (x1.001 / x0.999)
 *   - new_adjust_centroids . CodeVector/codevector_functions.py:414-439 (13-dim mean, zeros if empty)
 *   - createCodeVector ..... CodeVector/codevector_functions.py:442-531
 *
 * Arithmetic pin: np.linalg.norm(d) for a 12-vector is sqrt(d.dot(d)); with the numpy
 * 2.3.5 / OpenBLAS 0.3.30 build in the image the dot is a SEQUENTIAL FMA chain
 * acc = fma(d_i, d_i, acc), i = 0..11 (checked bit-for-bit on 20 000 random pairs,
 * see DESIGN.md "VQ arithmetic").  This file and the CUDA kernel use that chain, so
 * indices are bit-exact with the reference as run in the build container.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define DIM 13

static inline double dist2_sw(const double *x, const double *c) {
    double acc = 0.0;
    for (int d = 1; d < DIM; ++d) {
        double v = x[d] - c[d];
        acc = fma(v, v, acc);
    }
    return acc;
}

#if defined(__x86_64__)
__attribute__((target("fma"))) static inline double dist2_hw(const double *x, const double *c) {
    double acc = 0.0;
    for (int d = 1; d < DIM; ++d) {
        double v = x[d] - c[d];
        acc = __builtin_fma(v, v, acc);
    }
    return acc;
}
static int have_fma(void) { return __builtin_cpu_supports("fma"); }
#else
#define dist2_hw dist2_sw
static int have_fma(void) { return 0; }
#endif

/* One frame against K centroids: HMM/hmm_training.py:102-116 (and the identical loop at
 * codevector_functions.py:491-500).  Returns the winning index, *dist = its distance. */
#define ASSIGN_BODY(DIST2)                                              \
    double min_d = INFINITY;                                            \
    int best = 0;                                                       \
    for (int k = 0; k < K; ++k) {                                       \
        double dd = sqrt(DIST2(x, C + (size_t)k * DIM));                \
        if (dd < min_d) { min_d = dd; best = k; }                       \
    }                                                                   \
    *dist = min_d;                                                      \
    return best;

static int assign_sw(const double *x, const double *C, int K, double *dist) { ASSIGN_BODY(dist2_sw) }
#if defined(__x86_64__)
__attribute__((target("fma")))
#endif
static int assign_hw(const double *x, const double *C, int K, double *dist) { ASSIGN_BODY(dist2_hw) }

static void assign_all(const double *X, long F, const double *C, int K, int *idx, double *dist) {
    const int hw = have_fma();
#pragma omp parallel for schedule(static)
    for (long f = 0; f < F; ++f) {
        double d;
        idx[f] = hw ? assign_hw(X + (size_t)f * DIM, C, K, &d) : assign_sw(X + (size_t)f * DIM, C, K, &d);
        dist[f] = d;
    }
}

/* get_observations for a flat [F,13] frame matrix.  dist_out may be NULL. */
int vqo_encode(const double *X, long F, const double *C, int K, int *idx_out, double *dist_out) {
    if (F < 0 || K <= 0) return -1;
    double *dist = dist_out ? dist_out : (double *)malloc(sizeof(double) * (size_t)(F > 0 ? F : 1));
    if (!dist) return -2;
    assign_all(X, F, C, K, idx_out, dist);
    if (!dist_out) free(dist);
    return 0;
}

/* createCodeVector.  K = centroids_quantity; n_gen = int(log2(K)).  Outputs:
 *   C_out      [2^max(n_gen,1), 13]  final centroids (for K == 1 the reference returns the
 *                                    un-refined split pair, :469 + :531)
 *   gens_out   concatenated generations: [C0] then the converged 2^g centroids of each
 *              generation g = 1..n_gen  (1 + 2 + 4 + ... rows of 13)
 *   assign_out [F] frame.parent_centroid_id after the last assignment pass
 *   iters_out  [n_gen] Lloyd iterations used per generation
 *   gdist_out  [n_gen] last global distance per generation (may be NULL)
 */
int vqo_lbg(const double *X, long F, int K, int max_iter, double eps, double *C_out, double *gens_out,
            int *assign_out, int *iters_out, double *gdist_out) {
    if (F <= 0) return -1; /* ValueError("No raw data provided") :445-446 */
    if (K <= 0) return -3;
    int n_gen = (int)floor(log2((double)K));
    int Kmax = 1 << (n_gen > 0 ? n_gen : 1);
    double *cur = (double *)calloc((size_t)Kmax * DIM, sizeof(double));
    double *nxt = (double *)calloc((size_t)Kmax * DIM, sizeof(double));
    double *dist = (double *)malloc(sizeof(double) * (size_t)F);
    long *cnt = (long *)malloc(sizeof(long) * (size_t)Kmax);
    if (!cur || !nxt || !dist || !cnt) return -2;

    /* C0 = np.mean(all_mfcc, axis=0)  :458-459 (row-by-row accumulation) */
    double c0[DIM] = {0};
    for (long f = 0; f < F; ++f)
        for (int d = 0; d < DIM; ++d) c0[d] += X[(size_t)f * DIM + d];
    for (int d = 0; d < DIM; ++d) c0[d] /= (double)F;
    size_t gpos = 0;
    memcpy(gens_out + gpos, c0, sizeof(c0));
    gpos += DIM;
    /* first split :469 */
    int Kg = 2;
    for (int d = 0; d < DIM; ++d) { cur[d] = c0[d] * 1.001; cur[DIM + d] = c0[d] * 0.999; }
    for (long f = 0; f < F; ++f) assign_out[f] = 0;

    for (int g = 1; g <= n_gen; ++g) {
        double prev = 0.0, diff = eps + 100.0, gd = 0.0; /* :475-476 */
        int it = 0;
        while (diff > eps && it < max_iter) { /* :485 */
            ++it;
            assign_all(X, F, cur, Kg, assign_out, dist); /* :490-502 */
            gd = 0.0;
            for (long f = 0; f < F; ++f) gd += dist[f]; /* :503, frame order */
            /* new_adjust_centroids :414-439 */
            memset(nxt, 0, sizeof(double) * (size_t)Kg * DIM);
            memset(cnt, 0, sizeof(long) * (size_t)Kg);
            for (long f = 0; f < F; ++f) {
                int k = assign_out[f];
                cnt[k]++;
                for (int d = 0; d < DIM; ++d) nxt[(size_t)k * DIM + d] += X[(size_t)f * DIM + d];
            }
            for (int k = 0; k < Kg; ++k)
                if (cnt[k] > 0)
                    for (int d = 0; d < DIM; ++d) nxt[(size_t)k * DIM + d] /= (double)cnt[k];
            double *t = cur; cur = nxt; nxt = t;
            diff = fabs(prev - gd); /* :509-510 */
            prev = gd;
        }
        iters_out[g - 1] = it;
        if (gdist_out) gdist_out[g - 1] = gd;
        memcpy(gens_out + gpos, cur, sizeof(double) * (size_t)Kg * DIM); /* :517 */
        gpos += (size_t)Kg * DIM;
        if (g < n_gen) { /* :520-521 */
            for (int k = Kg - 1; k >= 0; --k)
                for (int d = 0; d < DIM; ++d) {
                    double v = cur[(size_t)k * DIM + d];
                    nxt[(size_t)(2 * k) * DIM + d] = v * 1.001;
                    nxt[(size_t)(2 * k + 1) * DIM + d] = v * 0.999;
                }
            double *t = cur; cur = nxt; nxt = t;
            Kg *= 2;
        }
    }
    memcpy(C_out, cur, sizeof(double) * (size_t)Kg * DIM);
    free(cur); free(nxt); free(dist); free(cnt);
    return Kg;
}
