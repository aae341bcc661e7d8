/* Plain C caller of the C ABI (include/hmmb200.h): no Python, no torch types anywhere on this side.
 *
 *   gcc -std=c11 -O2 -Iinclude examples/c_caller.c -Lhmm_training_b200 -lhmmb200 -lm -o c_caller
 *   LD_LIBRARY_PATH=hmm_training_b200 ./c_caller
 *
 * Trains two 4-state word models on synthetic codeword sequences (hmmb_bw_fit = hmm_training,
 * HMM/hmm_training.py:265-541, for all words at once), then scores the training utterances against both models
 * (hmmb_score = calculate_log_likelihood + the argmax of test_hmm, HMM/hmm_testing.py:49-104,139-161) and VQ-encodes
 * a few frames (hmmb_vq_encode = get_observations, HMM/hmm_training.py:82-120).  Exit code 0 = every utterance was
 * recognised as its own word. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hmmb200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ < 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, hmmb_last_error()); \
            return 1;                                                                \
        }                                                                            \
    } while (0)

enum { W = 2, N = 4, M = 16, S = 8, T = 24, R = W * S, ITERS = 5 };

static unsigned lcg(unsigned *s) { return *s = *s * 1664525u + 1013904223u; }

int main(void) {
    /* word w emits, in its k-th quarter, codewords from {4k + 2w, 4k + 2w + 1} mod M: left-to-right structure */
    static uint8_t obs[R * T];
    int64_t offsets[R + 1];
    int32_t word_of_seq[R];
    unsigned seed = 12345u;
    for (int r = 0; r < R; ++r) {
        const int w = r / S;
        word_of_seq[r] = w;
        offsets[r] = (int64_t)r * T;
        for (int t = 0; t < T; ++t) obs[r * T + t] = (uint8_t)((4 * (t * N / T) + 2 * w + (int)(lcg(&seed) >> 31)) % M);
    }
    offsets[R] = (int64_t)R * T;

    /* the reference's default initial model (hmm_training.py:300-320), one copy per word */
    static double pi0[W * N], A0[W * N * N], B0[W * N * M];
    for (int w = 0; w < W; ++w) {
        const double p[N] = {0.97, 0.02, 0.005, 0.005};
        memcpy(pi0 + w * N, p, sizeof p);
        for (int i = 0; i < N; ++i) {
            A0[(w * N + i) * N + i] = i + 1 < N ? 0.6 : 1.0;
            if (i + 1 < N) A0[(w * N + i) * N + i + 1] = 0.4;
        }
        for (int e = 0; e < N * M; ++e) B0[w * N * M + e] = 1.0 / M;
    }

    CHECK(hmmb_init(0));
    printf("%s\n", hmmb_version());

    static double pi[W * N], A[W * N * N], B[W * N * M], ll_hist[W * ITERS];
    int32_t iters[W];
    CHECK(hmmb_bw_fit(obs, 1, offsets, word_of_seq, R, W, N, M, pi0, A0, B0, 1e-6, ITERS, pi, A, B, ll_hist, iters));
    for (int w = 0; w < W; ++w) {
        printf("word %d: %d iterations, statistic %.6f -> %.6f\n", w, iters[w], ll_hist[w * ITERS], ll_hist[w * ITERS + iters[w] - 1]);
        if (!(ll_hist[w * ITERS + iters[w] - 1] >= ll_hist[w * ITERS] - 1e-9)) return 2; /* EM does not decrease it */
        for (int i = 0; i < N; ++i) {
            double row = 0.0;
            for (int k = 0; k < M; ++k) row += B[(w * N + i) * M + k];
            if (fabs(row - 1.0) > 1e-9) return 3; /* exit normalisation, hmm_training.py:524-539 */
        }
    }

    static double ll[R * W];
    int32_t best[R];
    CHECK(hmmb_score(obs, 1, 0, offsets, R, W, N, M, pi, A, B, ll, best));
    int wrong = 0;
    for (int r = 0; r < R; ++r) wrong += best[r] != word_of_seq[r];
    printf("recognition: %d of %d utterances assigned to their own word\n", R - wrong, R);

    /* VQ: every centroid encodes to itself (dimension 0, the energy, is ignored: hmm_training.py:100,107) */
    static double C[4 * 13], X[4 * 13];
    int32_t idx[4];
    for (int k = 0; k < 4; ++k)
        for (int d = 0; d < 13; ++d) {
            C[k * 13 + d] = (double)(k * 13 + d) * (k % 2 ? -1.0 : 1.0);
            X[k * 13 + d] = d == 0 ? 1e6 : C[k * 13 + d];
        }
    CHECK(hmmb_vq_encode(X, 4, C, 4, idx));
    for (int k = 0; k < 4; ++k) wrong += idx[k] != k;

    CHECK(hmmb_shutdown());
    return wrong ? 4 : 0;
}
