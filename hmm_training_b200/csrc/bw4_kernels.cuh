// N = 4 Baum-Welch E-step kernels (the reference's hard-coded state count,
// HMM/hmm_training.py:226): one sequence per thread, 32 same-word sequences per warp in
// lock-step.
//
//   k_repack_blocks4  ragged codewords -> blocked u16 layout [block][chunk of 8 steps][lane];
//                     each entry = codeword (11 bits) | conflict rank (5 bits): the number of
//                     lower lanes of the warp that see the same codeword at the same step.
//                     Codewords never change during training, so the conflict schedule of
//                     the emission-count update is computed once here instead of with
//                     MATCH.ANY every step (measured 1 MATCH per ~55 cycles per SM on B200).
//   k_bw_fwd4         scaled forward pass, alpha-hat spilled to HBM [block][t][half][lane] as
//                     coalesced 16-byte streaming stores, B^T of the CTA's word in shared memory.
//   k_bw_bwd4         backward pass fused with the gamma / xi / emission-count accumulation
//                     (beta never leaves registers); counts go to warp-private shared-memory
//                     copies with plain read-modify-writes in precomputed rank order (no
//                     atomics, deterministic); one partial per CTA.
//
// BIDIAG = true specialises both kernels for an upper-bidiagonal transition matrix (the
// reference's left-to-right default, :307-312 — zeros of A stay zeros under re-estimation,
// :450-455): 7 instead of 16 products in every mat-vec and 7 instead of 16 xi accumulators.
// Skipped terms are exact zeros, so the results are bit-identical to the dense kernel.
#pragma once

#include "hmm_device.cuh"

namespace hmmb {

constexpr int BW_THREADS = 128;  // 4 warps per CTA in every E-step kernel
constexpr int BW_WARPS = BW_THREADS / 32;
constexpr int SPC4 = 8;          // packed u16 entries per uint4
constexpr unsigned SYM_MASK = 0x7ffu;
constexpr int SYM_BITS = 11;
constexpr int BW4_MAX_M = 512;   // warp-private count copies must fit in shared memory

// ---------------------------------------------------------------- repack
template <typename InT>
__global__ void __launch_bounds__(256)
k_repack_blocks4(const InT *__restrict__ obs, const int64_t *__restrict__ off_sorted,
                 const int32_t *__restrict__ len_sorted, const Blk *__restrict__ blks, int nblk,
                 uint4 *__restrict__ obs_blk, int M, int *__restrict__ bad) {
    __shared__ unsigned short sSym[SPC4][32];
    const int b = blockIdx.x;
    if (b >= nblk) return;
    const Blk bk = blks[b];
    const int nch = (bk.tmax + SPC4 - 1) / SPC4;
    const int tid = threadIdx.x;  // 256 threads = 8 steps x 32 lanes
    const int s = tid >> 5, lane = tid & 31;
    int T = 0;
    const InT *src = obs;
    if (lane < bk.nseq) {
        T = len_sorted[bk.first + lane];
        src = obs + off_sorted[bk.first + lane];
    }
    for (int c = 0; c < nch; ++c) {
        const int t = c * SPC4 + s;
        unsigned sym = 0xffffu;  // marks "no frame"
        if (t < T) {
            unsigned long long v = (unsigned long long)src[t];
            if (v >= (unsigned long long)M) { atomicOr(bad, 1); v = 0; }
            sym = (unsigned)v;
        }
        __syncthreads();
        sSym[s][lane] = (unsigned short)sym;
        __syncthreads();
        unsigned rank = 0;
        if (sym != 0xffffu)
            for (int l = 0; l < lane; ++l) rank += (sSym[s][l] == sym) ? 1u : 0u;
        const unsigned packed = (sym == 0xffffu) ? 0u : (sym | (rank << SYM_BITS));
        __syncthreads();
        sSym[s][lane] = (unsigned short)packed;
        __syncthreads();
        if (s == 0) {
            unsigned r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = (unsigned)sSym[2 * q][lane] | ((unsigned)sSym[2 * q + 1][lane] << 16);
            obs_blk[bk.obs_base + (size_t)c * 32 + lane] = make_uint4(r[0], r[1], r[2], r[3]);
        }
    }
}

// ---------------------------------------------------------------- forward
// zero-or-denormal test for four non-negative doubles at once (integer pipe)
__device__ __forceinline__ bool any_sub4(double x0, double x1, double x2, double x3) {
    const unsigned m = min(min((unsigned)__double2hiint(x0), (unsigned)__double2hiint(x1)),
                           min((unsigned)__double2hiint(x2), (unsigned)__double2hiint(x3)));
    return m < 0x00100000u;
}

// a[] holds A row-major (dense) or {a00,a11,a22,a33,a01,a12,a23} (BIDIAG).
template <bool BIDIAG>
__device__ __forceinline__ void matvec_fwd(const double *a, double al0, double al1, double al2, double al3, double &n0,
                                           double &n1, double &n2, double &n3) {
    if (BIDIAG) {
        n0 = al0 * a[0];
        n1 = al0 * a[4] + al1 * a[1];
        n2 = al1 * a[5] + al2 * a[2];
        n3 = al2 * a[6] + al3 * a[3];
    } else {
        n0 = al0 * a[0] + al1 * a[4] + al2 * a[8] + al3 * a[12];
        n1 = al0 * a[1] + al1 * a[5] + al2 * a[9] + al3 * a[13];
        n2 = al0 * a[2] + al1 * a[6] + al2 * a[10] + al3 * a[14];
        n3 = al0 * a[3] + al1 * a[7] + al2 * a[11] + al3 * a[15];
    }
}
// A[i][j] from either representation (slow paths only)
template <bool BIDIAG>
__device__ __forceinline__ double a_at(const double *a, int i, int j) {
    if (BIDIAG) return (j == i) ? a[i] : ((j == i + 1) ? a[4 + i] : 0.0);
    return a[i * 4 + j];
}

// One sequence per lane.  p[j] = pi[j], sB[sym*4+j] = B[j][sym] (shared), sBmax[sym] =
// max_j B[j][sym], rmax = largest row sum of A.  op / sp include the lane offset.  Replaces
// calculate_log_alpha (HMM/hmm_training.py:122-160) and the alpha init (:357-360); returns
// log P(O|lambda) (:376-377), -inf for a structurally impossible sequence, or NaN when the
// precision guard asks for the exact log-space recomputation.
//
// Precision guard, N = 4 flavour: every alive state whose value is denormal / clamped adds
// 2^-1074 (its worst-case absolute error) to a scalar bound E, which is propagated with
//   E' <= (E * rmax * max_j b_j(o_t) + seeds) * scale   >=  sum_j |error of alpha_t(j)|
// (units of 2^-1000).  E stays ~1e-320 relative unless the states that carried the sequence
// die; when it exceeds 1e-12 of the step's mass the sequence is handed over.
template <bool BIDIAG, bool SPILL>
__device__ __forceinline__ double fwd4_run(int T, int tmax, const uint4 *__restrict__ op,
                                           const double *__restrict__ sB, const double *__restrict__ sBmax,
                                           const double *a, const double (&p)[4], double rmax,
                                           double2 *__restrict__ sp) {
    using S16 = Sym<uint16_t>;
    double al0 = 0.0, al1 = 0.0, al2 = 0.0, al3 = 0.0;
    double E = 0.0;  // error bound, units of 2^-1000
    long long esum = 0;
    bool stop = false;  // dead (impossible) or flagged for the exact path
    double ll = neg_inf();
    const int nch = (tmax + SPC4 - 1) / SPC4;
    uint4 wnext = nch > 0 ? __ldg(op) : make_uint4(0, 0, 0, 0);
    for (int c = 0; c < nch; ++c) {
        uint4 w = wnext;
        if (c + 1 < nch) wnext = __ldg(op + (size_t)(c + 1) * 32);  // prefetch the next 8 codewords
#pragma unroll 2
        for (int s = 0; s < SPC4; ++s) {
            const int t = c * SPC4 + s;
            const unsigned sym = S16::pop_front(w) & SYM_MASK;
            if (t < T && !stop) {
                const double2 b01 = *reinterpret_cast<const double2 *>(sB + sym * 4);
                const double2 b23 = *reinterpret_cast<const double2 *>(sB + sym * 4 + 2);
                double n0, n1, n2, n3;
                if (t == 0) {
                    n0 = p[0]; n1 = p[1]; n2 = p[2]; n3 = p[3];
                } else {
                    matvec_fwd<BIDIAG>(a, al0, al1, al2, al3, n0, n1, n2, n3);
                }
                double at0 = n0 * b01.x, at1 = n1 * b01.y, at2 = n2 * b23.x, at3 = n3 * b23.y;
                double ssum = (at0 + at1) + (at2 + at3);
                double seeds = 0.0;
                if (!(ssum >= TINY_STEP)) {
                    // ---- the whole step is tiny (or impossible): exponent-split products
#define HMMB_FIX_N(J, NJ)                                                                                      \
    if (NJ == 0.0 && t > 0 &&                                                                                  \
        ((al0 > 0.0 && a_at<BIDIAG>(a, 0, J) > 0.0) || (al1 > 0.0 && a_at<BIDIAG>(a, 1, J) > 0.0) ||           \
         (al2 > 0.0 && a_at<BIDIAG>(a, 2, J) > 0.0) || (al3 > 0.0 && a_at<BIDIAG>(a, 3, J) > 0.0)))            \
        NJ = tiny_pos();
                    HMMB_FIX_N(0, n0) HMMB_FIX_N(1, n1) HMMB_FIX_N(2, n2) HMMB_FIX_N(3, n3)
                    double o[4];
                    int Ex;
                    const int code = exact_products4(n0, n1, n2, n3, b01.x, b01.y, b23.x, b23.y, o, &Ex);
                    if (code == 0) {
                        stop = true;  // no state can emit o_t: log P = -inf
                    } else if ((code == 2 && t > 0) || E > 0.0) {
                        stop = true;  // the surviving states had lost their bits: exact path
                        ll = nan_mark();
                    } else {
                        esum += Ex;
                        at0 = o[0]; at1 = o[1]; at2 = o[2]; at3 = o[3];
                        ssum = (at0 + at1) + (at2 + at3);
                    }
                } else if (any_sub4(at0, at1, at2, at3)) {
                    // ---- some state is zero / denormal: keep "alpha_j > 0 <=> structurally
                    // reachable" (clamp) and count the seeds of the error bound
                    HMMB_FIX_N(0, n0) HMMB_FIX_N(1, n1) HMMB_FIX_N(2, n2) HMMB_FIX_N(3, n3)
#undef HMMB_FIX_N
                    if (n0 > 0.0 && b01.x > 0.0 && is_sub(at0)) { if (at0 == 0.0) at0 = tiny_pos(); seeds += ERR_UNIT; }
                    if (n1 > 0.0 && b01.y > 0.0 && is_sub(at1)) { if (at1 == 0.0) at1 = tiny_pos(); seeds += ERR_UNIT; }
                    if (n2 > 0.0 && b23.x > 0.0 && is_sub(at2)) { if (at2 == 0.0) at2 = tiny_pos(); seeds += ERR_UNIT; }
                    if (n3 > 0.0 && b23.y > 0.0 && is_sub(at3)) { if (at3 == 0.0) at3 = tiny_pos(); seeds += ERR_UNIT; }
                }
                if (!stop) {
                    const double sc = pow2_rescale(ssum, esum);
                    al0 = at0 * sc; al1 = at1 * sc; al2 = at2 * sc; al3 = at3 * sc;
                    if ((E > 0.0) | (seeds > 0.0)) {
                        E = (E * (rmax * sBmax[sym]) + seeds) * sc;
                        if (!(E <= ERR_LIMIT)) {
                            stop = true;
                            ll = nan_mark();
                        }
                    }
                    if (t == T - 1 && !stop) ll = log((al0 + al1) + (al2 + al3)) + (double)esum * LN2;
                }
                if (stop) al0 = al1 = al2 = al3 = 0.0;
                if (SPILL) {
                    __stcs(sp + (size_t)t * 64, make_double2(al0, al1));
                    __stcs(sp + (size_t)t * 64 + 32, make_double2(al2, al3));
                }
            }
        }
    }
    return ll;
}

// CTA prologue shared by the forward-type kernels: B^T of word w -> shared memory, per-codeword
// max_j b_j, A and pi -> registers, rmax = largest row sum of A.
template <bool BIDIAG>
__device__ __forceinline__ void load_model4(const double *__restrict__ pi, const double *__restrict__ A,
                                            const double *__restrict__ Bt, int w, int M, double *sB, double *sBmax,
                                            double *a, double (&p)[4], double &rmax) {
    const int tid = threadIdx.x;
    const double2 *src = reinterpret_cast<const double2 *>(Bt + (size_t)w * M * 4);
    double2 *dst = reinterpret_cast<double2 *>(sB);
    for (int e = tid; e < M; e += BW_THREADS) {
        const double2 x = __ldg(src + 2 * e), y = __ldg(src + 2 * e + 1);
        dst[2 * e] = x;
        dst[2 * e + 1] = y;
        sBmax[e] = fmax(fmax(x.x, x.y), fmax(y.x, y.y));
    }
    const double *Aw = A + (size_t)w * 16;
    rmax = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        rmax = fmax(rmax, (__ldg(Aw + i * 4) + __ldg(Aw + i * 4 + 1)) + (__ldg(Aw + i * 4 + 2) + __ldg(Aw + i * 4 + 3)));
    if (BIDIAG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = __ldg(Aw + i * 5);
#pragma unroll
        for (int i = 0; i < 3; ++i) a[4 + i] = __ldg(Aw + i * 5 + 1);
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] = __ldg(Aw + q);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) p[q] = __ldg(pi + (size_t)w * 4 + q);
}

template <bool BIDIAG>
__device__ __forceinline__ void load_A4(const double *__restrict__ Aw, double *a) {
    if (BIDIAG) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = __ldg(Aw + i * 5);
#pragma unroll
        for (int i = 0; i < 3; ++i) a[4 + i] = __ldg(Aw + i * 5 + 1);
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) a[q] = __ldg(Aw + q);
    }
}

template <bool BIDIAG>
__global__ void __launch_bounds__(BW_THREADS)
k_bw_fwd4(const CtaWork *__restrict__ work, const Blk *__restrict__ blks, const uint4 *__restrict__ obs_blk,
          const int32_t *__restrict__ len_sorted, const double *__restrict__ pi, const double *__restrict__ A,
          const double *__restrict__ Bt, int M, double2 *__restrict__ spill, double *__restrict__ ll_seq,
          const int32_t *__restrict__ active, uint8_t *__restrict__ flag) {
    extern __shared__ double sB[];  // [M][4] B^T, then [M] per-codeword max
    double *sBmax = sB + (size_t)M * 4;
    const CtaWork cw = work[blockIdx.x];
    if (!active[cw.word]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double a[BIDIAG ? 7 : 16], p[4], rmax;
    load_model4<BIDIAG>(pi, A, Bt, cw.word, M, sB, sBmax, a, p, rmax);
    __syncthreads();
    for (int b = cw.blk_begin + warp; b < cw.blk_end; b += BW_WARPS) {
        const Blk bk = blks[b];
        int T = lane < bk.nseq ? len_sorted[bk.first + lane] : 0;
        if (T > 0 && flag[bk.first + lane]) T = 0;  // handled by the exact log-space kernel
        const double ll = fwd4_run<BIDIAG, true>(T, bk.tmax, obs_blk + bk.obs_base + lane, sB, sBmax, a, p, rmax,
                                                 spill + bk.spill_base * 64 + lane);
        if (T > 0) {
            ll_seq[bk.first + lane] = ll;
            if (ll != ll) flag[bk.first + lane] = 1;  // precision guard: hand over (sticky)
        }
    }
}

// ---------------------------------------------------------------- backward + accumulate
// Warp-private emission-count update in precomputed rank order (see k_repack_blocks4).  Lanes
// of rank 0 (the common case) have distinct codewords; their read-modify-write is split
// around the next step's arithmetic by the caller so the shared-memory latency is hidden.
// Ranks >= 1 follow here, one conflict-free round per rank.
__device__ __forceinline__ void cnt_update4_rest(double *__restrict__ cw, bool act, unsigned sym, int rank, double g0,
                                                 double g1, double g2, double g3) {
    const int maxrank = __reduce_max_sync(0xffffffffu, act ? rank : 0);
    double2 *row = reinterpret_cast<double2 *>(cw + sym * 4);
    for (int r = 1; r <= maxrank; ++r) {
        if (act && rank == r) {
            double2 c01 = row[0], c23 = row[1];
            c01.x += g0; c01.y += g1; c23.x += g2; c23.y += g3;
            row[0] = c01; row[1] = c23;
        }
        __syncwarp();
    }
}

// all four non-negative doubles strictly positive? (integer pipe: x > 0 <=> hi|lo != 0)
__device__ __forceinline__ bool all_pos4(double x0, double x1, double x2, double x3) {
    const unsigned m0 = (unsigned)__double2hiint(x0) | (unsigned)__double2loint(x0);
    const unsigned m1 = (unsigned)__double2hiint(x1) | (unsigned)__double2loint(x1);
    const unsigned m2 = (unsigned)__double2hiint(x2) | (unsigned)__double2loint(x2);
    const unsigned m3 = (unsigned)__double2hiint(x3) | (unsigned)__double2loint(x3);
    return min(min(m0, m1), min(m2, m3)) != 0u;
}
// x == 0 -> smallest denormal, else x (non-negative input; integer pipe, branch-free)
__device__ __forceinline__ double zero_to_tiny(double x) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    return __hiloint2double(hi, lo | (((hi | lo) == 0) ? 1 : 0));
}
constexpr double LEAN_MIN = 0x1p-500;  // the lean backward step needs its three sums above this

// State of one lane's backward recursion.
template <bool BIDIAG>
struct Bwd4State {
    double v0, v1, v2, v3;        // v_j = b_j(o_{t+1}) * beta-hat_{t+1}(j)
    double X[BIDIAG ? 7 : 16];    // sum_t u_i w_j (a_ij applied at the flush)
    unsigned seenX;               // (i,j) pairs for which a finite xi term existed
    bool imprecise;
    bool vpos;                    // every v_j > 0 (lets the lean path skip the structural masks)
};

// Careful version of one backward step (all clamps, structural masks, precision hand-over).
// Taken when the lean path's two group tests fail, at the first step of a sequence, and for
// words whose B holds exact zeros.  Returns gamma in g[], updates st (v, X, seenX).
template <bool BIDIAG>
__device__ __noinline__ void bwd4_step_slow(Bwd4State<BIDIAG> &st, const double *a, const double *__restrict__ sB,
                                            unsigned sym, bool last, double al0, double al1, double al2, double al3,
                                            double *g) {
    double v0 = st.v0, v1 = st.v1, v2 = st.v2, v3 = st.v3;
    double h0 = 1.0, h1 = 1.0, h2 = 1.0, h3 = 1.0;  // log beta_{T-1} = 0 (:363)
    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
    if (!last) {
        // un-normalised beta_t(i) = sum_j a_ij b_j(o_{t+1}) beta_{t+1}(j)  (:163-199)
        double q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            q[i] = a_at<BIDIAG>(a, i, 0) * v0 + a_at<BIDIAG>(a, i, 1) * v1 + a_at<BIDIAG>(a, i, 2) * v2 +
                   a_at<BIDIAG>(a, i, 3) * v3;
            if (q[i] == 0.0 && ((a_at<BIDIAG>(a, i, 0) > 0.0 && v0 > 0.0) || (a_at<BIDIAG>(a, i, 1) > 0.0 && v1 > 0.0) ||
                                (a_at<BIDIAG>(a, i, 2) > 0.0 && v2 > 0.0) || (a_at<BIDIAG>(a, i, 3) > 0.0 && v3 > 0.0)))
                q[i] = tiny_pos();
        }
        const double qs = (q[0] + q[1]) + (q[2] + q[3]);
        const double sc = qs > 0.0 ? pow2_rescale_noacc(qs) : 1.0;
        h0 = q[0] * sc; h1 = q[1] * sc; h2 = q[2] * sc; h3 = q[3] * sc;
        w0 = v0 * sc; w1 = v1 * sc; w2 = v2 * sc; w3 = v3 * sc;
    }
    // gamma_t(i) = alpha_t(i) beta_t(i) / sum_i alpha_t(i) beta_t(i)   (:389-394)
    double g0 = al0 * h0, g1 = al1 * h1, g2 = al2 * h2, g3 = al3 * h3;
    double norm = (g0 + g1) + (g2 + g3);
    double u0 = al0, u1 = al1, u2 = al2, u3 = al3;
    double r;
    if (!(norm >= TINY_STEP)) {
        // forward and backward mass sit on (almost) disjoint states: redo the products 2^1000
        // larger so they neither underflow nor blow up 1/norm; the bits may already be gone,
        // so the sequence is also handed over to the exact kernel
        st.imprecise = true;
        const double big = 0x1p500;
        u0 = al0 * big; u1 = al1 * big; u2 = al2 * big; u3 = al3 * big;
        g0 = u0 * (h0 * big); g1 = u1 * (h1 * big); g2 = u2 * (h2 * big); g3 = u3 * (h3 * big);
        w0 *= big; w1 *= big; w2 *= big; w3 *= big;
        norm = (g0 + g1) + (g2 + g3);
        r = norm > 0.0 ? 1.0 / norm : 0.0;
    } else {
        r = 1.0 / norm;
    }
    g0 *= r; g1 *= r; g2 *= r; g3 *= r;
    if (g0 == 0.0 && al0 > 0.0 && h0 > 0.0) g0 = tiny_pos();
    if (g1 == 0.0 && al1 > 0.0 && h1 > 0.0) g1 = tiny_pos();
    if (g2 == 0.0 && al2 > 0.0 && h2 > 0.0) g2 = tiny_pos();
    if (g3 == 0.0 && al3 > 0.0 && h3 > 0.0) g3 = tiny_pos();
    g[0] = g0; g[1] = g1; g[2] = g2; g[3] = g3;
    if (!last) {
        // xi_t(i,j) = alpha_t(i) a_ij b_j(o_{t+1}) beta_{t+1}(j) / norm   (:397-410)
        u0 *= r; u1 *= r; u2 *= r; u3 *= r;
        const double uu[4] = {u0, u1, u2, u3}, ww[4] = {w0, w1, w2, w3};
        if (BIDIAG) {
#pragma unroll
            for (int i = 0; i < 4; ++i) st.X[i] = fma(uu[i], ww[i], st.X[i]);
#pragma unroll
            for (int i = 0; i < 3; ++i) st.X[4 + i] = fma(uu[i], ww[i + 1], st.X[4 + i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) st.X[i * 4 + j] = fma(uu[i], ww[j], st.X[i * 4 + j]);
        }
        const unsigned mv = (v0 > 0.0 ? 1u : 0u) | (v1 > 0.0 ? 2u : 0u) | (v2 > 0.0 ? 4u : 0u) | (v3 > 0.0 ? 8u : 0u);
        st.seenX |= (al0 > 0.0 ? mv : 0u) | (al1 > 0.0 ? mv << 4 : 0u) | (al2 > 0.0 ? mv << 8 : 0u) |
                    (al3 > 0.0 ? mv << 12 : 0u);
    }
    // v_j = b_j(o_t) beta-hat_t(j) for step t-1
    const double2 b01 = *reinterpret_cast<const double2 *>(sB + sym * 4);
    const double2 b23 = *reinterpret_cast<const double2 *>(sB + sym * 4 + 2);
    v0 = b01.x * h0; v1 = b01.y * h1; v2 = b23.x * h2; v3 = b23.y * h3;
    const double vs = (v0 + v1) + (v2 + v3);
    if (!(vs >= TINY_STEP)) {
        // tiny emission column: exponent-split products (the scale of v is free)
        double o[4];
        int E;
        if (exact_products4(h0, h1, h2, h3, b01.x, b01.y, b23.x, b23.y, o, &E) == 2) st.imprecise = true;
        v0 = o[0]; v1 = o[1]; v2 = o[2]; v3 = o[3];
    } else {
        if (v0 == 0.0 && b01.x > 0.0 && h0 > 0.0) v0 = tiny_pos();
        if (v1 == 0.0 && b01.y > 0.0 && h1 > 0.0) v1 = tiny_pos();
        if (v2 == 0.0 && b23.x > 0.0 && h2 > 0.0) v2 = tiny_pos();
        if (v3 == 0.0 && b23.y > 0.0 && h3 > 0.0) v3 = tiny_pos();
    }
    st.v0 = v0; st.v1 = v1; st.v2 = v2; st.v3 = v3;
    st.vpos = (v0 > 0.0) & (v1 > 0.0) & (v2 > 0.0) & (v3 > 0.0);
}

// Partial layout per CTA (and accumulator layout per word): [pi N][xi N*N][cnt M*N].
template <bool BIDIAG>
__global__ void __launch_bounds__(BW_THREADS, BIDIAG ? 4 : 3)
k_bw_bwd4(const CtaWork *__restrict__ work, const Blk *__restrict__ blks, const uint4 *__restrict__ obs_blk,
          const int32_t *__restrict__ len_sorted, const double *__restrict__ A, const double *__restrict__ Bt, int M,
          const double2 *__restrict__ spill, const double *__restrict__ ll_seq, const int32_t *__restrict__ active,
          const int32_t *__restrict__ b_has_zero, double *__restrict__ partials, int64_t pstride,
          uint8_t *__restrict__ flag, int32_t *__restrict__ new_flags) {
    using S16 = Sym<uint16_t>;
    extern __shared__ double smem[];
    double *sB = smem;                              // [M][4]
    double *sCnt = smem + (size_t)M * 4;            // [4 warps][M][4]
    double *sPi = sCnt + (size_t)M * 4 * BW_WARPS;  // [128 threads][4] gamma_0 sums
    __shared__ double sRed[BW_WARPS][20];
    __shared__ unsigned sSeen;

    const CtaWork cw = work[blockIdx.x];
    double *part = partials + (size_t)blockIdx.x * pstride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!active[cw.word]) return;
    {
        const double2 *src = reinterpret_cast<const double2 *>(Bt + (size_t)cw.word * M * 4);
        double2 *dst = reinterpret_cast<double2 *>(sB);
        for (int e = tid; e < M * 2; e += BW_THREADS) dst[e] = __ldg(src + e);
        for (int e = tid; e < M * 4 * BW_WARPS + BW_THREADS * 4; e += BW_THREADS) sCnt[e] = 0.0;  // counts + sPi
        if (tid == 0) sSeen = 0u;
    }
    double a[BIDIAG ? 7 : 16];
    load_A4<BIDIAG>(A + (size_t)cw.word * 16, a);
    const bool lean_ok = b_has_zero[cw.word] == 0;  // B > 0 everywhere: v_j > 0 <=> beta_j > 0
    __syncthreads();

    double *cntw = sCnt + (size_t)warp * M * 4;
    double *mypi = sPi + (size_t)tid * 4;
    Bwd4State<BIDIAG> st;
#pragma unroll
    for (int q = 0; q < (BIDIAG ? 7 : 16); ++q) st.X[q] = 0.0;
    st.seenX = 0u;

    for (int b = cw.blk_begin + warp; b < cw.blk_end; b += BW_WARPS) {
        const Blk bk = blks[b];
        int T = 0;
        if (lane < bk.nseq) {
            T = len_sorted[bk.first + lane];
            if (!(ll_seq[bk.first + lane] > neg_inf())) T = 0;  // impossible sequence: contributes nothing (:391-394)
            if (flag[bk.first + lane]) T = 0;                   // exact log-space kernel did this one
        }
        st.v0 = st.v1 = st.v2 = st.v3 = 0.0;
        st.imprecise = false;
        st.vpos = false;
        const uint4 *op = obs_blk + bk.obs_base + lane;
        const double2 *sp = spill + bk.spill_base * 64 + lane;
        // alpha-hat prefetch, two steps deep: (pa*, t = tcur) and (pb*, t = tcur - 1)
        const int ttop = bk.tmax - 1;
        double2 pa01 = make_double2(0.0, 0.0), pa23 = pa01, pb01 = pa01, pb23 = pa01;
        if (ttop < T) { pa01 = __ldcs(sp + (size_t)ttop * 64); pa23 = __ldcs(sp + (size_t)ttop * 64 + 32); }
        if (ttop >= 1 && ttop - 1 < T) { pb01 = __ldcs(sp + (size_t)(ttop - 1) * 64); pb23 = __ldcs(sp + (size_t)(ttop - 1) * 64 + 32); }
        const int nch = (bk.tmax + SPC4 - 1) / SPC4;
        // emission-count update of the previous step, still pending (software pipelining)
        bool pact = false;
        unsigned psym = 0u;
        int prank = 0;
        double pg0 = 0.0, pg1 = 0.0, pg2 = 0.0, pg3 = 0.0;
        uint4 wnext = __ldg(op + (size_t)(nch - 1) * 32);
        for (int c = nch - 1; c >= 0; --c) {
            uint4 w = wnext;
            if (c > 0) wnext = __ldg(op + (size_t)(c - 1) * 32);  // prefetch the next 8 codewords
#pragma unroll 2
            for (int s = SPC4 - 1; s >= 0; --s) {
                const int t = c * SPC4 + s;
                const unsigned packed = S16::pop_back(w);
                if (t >= bk.tmax) continue;  // warp-uniform
                const unsigned sym = packed & SYM_MASK;
                const int rank = (int)(packed >> SYM_BITS);
                const bool act = t < T;
                // round 0 of the pending update: issue the shared-memory loads now ...
                const bool p0 = pact && prank == 0;
                double2 *prow = reinterpret_cast<double2 *>(cntw + psym * 4);
                double2 r01 = make_double2(0.0, 0.0), r23 = r01;
                if (p0) { r01 = prow[0]; r23 = prow[1]; }
                const double al0 = pa01.x, al1 = pa01.y, al2 = pa23.x, al3 = pa23.y;
                pa01 = pb01; pa23 = pb23;
                if (t >= 2 && t - 2 < T) {
                    pb01 = __ldcs(sp + (size_t)(t - 2) * 64);
                    pb23 = __ldcs(sp + (size_t)(t - 2) * 64 + 32);
                }
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                if (act) {
                    bool done = false;
                    if (lean_ok && st.vpos && t != T - 1) {
                        // ---- lean path: every state of the step alive, sums far from underflow.
                        // beta_t(i) ~ q_i = sum_j a_ij v_j  (:163-199);  gamma_t(i) = al_i q_i / norm
                        // (:389-394);  xi_t(i,j) = al_i a_ij v_j / norm  (:397-410), norm = sum_i al_i q_i
                        const double v0 = st.v0, v1 = st.v1, v2 = st.v2, v3 = st.v3;
                        double q0, q1, q2, q3;
                        if (BIDIAG) {
                            q0 = a[0] * v0 + a[4] * v1;
                            q1 = a[1] * v1 + a[5] * v2;
                            q2 = a[2] * v2 + a[6] * v3;
                            q3 = a[3] * v3;
                        } else {
                            q0 = a[0] * v0 + a[1] * v1 + a[2] * v2 + a[3] * v3;
                            q1 = a[4] * v0 + a[5] * v1 + a[6] * v2 + a[7] * v3;
                            q2 = a[8] * v0 + a[9] * v1 + a[10] * v2 + a[11] * v3;
                            q3 = a[12] * v0 + a[13] * v1 + a[14] * v2 + a[15] * v3;
                        }
                        const double qs = (q0 + q1) + (q2 + q3);
                        double c0 = al0 * q0, c1 = al1 * q1, c2 = al2 * q2, c3 = al3 * q3;
                        const double norm = (c0 + c1) + (c2 + c3);
                        const double sc = pow2_rescale_noacc(qs);
                        const double h0 = q0 * sc, h1 = q1 * sc, h2 = q2 * sc, h3 = q3 * sc;  // beta-hat_t
                        const double2 b01 = *reinterpret_cast<const double2 *>(sB + sym * 4);
                        const double2 b23 = *reinterpret_cast<const double2 *>(sB + sym * 4 + 2);
                        double nv0 = b01.x * h0, nv1 = b01.y * h1, nv2 = b23.x * h2, nv3 = b23.y * h3;
                        const double vs = (nv0 + nv1) + (nv2 + nv3);
                        if ((norm >= LEAN_MIN) & (qs >= LEAN_MIN) & (vs >= LEAN_MIN) & all_pos4(al0, al1, al2, al3) &
                            all_pos4(q0, q1, q2, q3)) {
                            const double r = 1.0 / norm;
                            // a finite log value that underflows stays (barely) positive
                            g0 = zero_to_tiny(c0 * r); g1 = zero_to_tiny(c1 * r);
                            g2 = zero_to_tiny(c2 * r); g3 = zero_to_tiny(c3 * r);
                            const double u0 = al0 * r, u1 = al1 * r, u2 = al2 * r, u3 = al3 * r;
                            if (BIDIAG) {
                                st.X[0] = fma(u0, v0, st.X[0]); st.X[1] = fma(u1, v1, st.X[1]);
                                st.X[2] = fma(u2, v2, st.X[2]); st.X[3] = fma(u3, v3, st.X[3]);
                                st.X[4] = fma(u0, v1, st.X[4]); st.X[5] = fma(u1, v2, st.X[5]);
                                st.X[6] = fma(u2, v3, st.X[6]);
                            } else {
                                st.X[0] = fma(u0, v0, st.X[0]);   st.X[1] = fma(u0, v1, st.X[1]);   st.X[2] = fma(u0, v2, st.X[2]);   st.X[3] = fma(u0, v3, st.X[3]);
                                st.X[4] = fma(u1, v0, st.X[4]);   st.X[5] = fma(u1, v1, st.X[5]);   st.X[6] = fma(u1, v2, st.X[6]);   st.X[7] = fma(u1, v3, st.X[7]);
                                st.X[8] = fma(u2, v0, st.X[8]);   st.X[9] = fma(u2, v1, st.X[9]);   st.X[10] = fma(u2, v2, st.X[10]); st.X[11] = fma(u2, v3, st.X[11]);
                                st.X[12] = fma(u3, v0, st.X[12]); st.X[13] = fma(u3, v1, st.X[13]); st.X[14] = fma(u3, v2, st.X[14]); st.X[15] = fma(u3, v3, st.X[15]);
                            }
                            st.seenX = 0xffffu;  // every alpha_i > 0 and (vpos) every v_j > 0
                            // B > 0 and beta-hat > 0: v stays positive (clamped if the product underflows)
                            st.v0 = zero_to_tiny(nv0); st.v1 = zero_to_tiny(nv1);
                            st.v2 = zero_to_tiny(nv2); st.v3 = zero_to_tiny(nv3);
                            done = true;
                        }
                    }
                    if (!done) {
                        // the careful step works on a stack copy so that `st` itself never has
                        // its address taken and stays in registers on the lean path
                        Bwd4State<BIDIAG> tmp = st;
                        double g[4];
                        bwd4_step_slow<BIDIAG>(tmp, a, sB, sym, t == T - 1, al0, al1, al2, al3, g);
                        st = tmp;
                        g0 = g[0]; g1 = g[1]; g2 = g[2]; g3 = g[3];
                    }
                    if (t == 0) {  // (:415-426)
                        mypi[0] += g0; mypi[1] += g1; mypi[2] += g2; mypi[3] += g3;
                    }
                }
                // ... and finish it after this step's arithmetic  (:460-500 numerators)
                if (p0) {
                    r01.x += pg0; r01.y += pg1; r23.x += pg2; r23.y += pg3;
                    prow[0] = r01; prow[1] = r23;
                }
                __syncwarp();
                cnt_update4_rest(cntw, pact, psym, prank, pg0, pg1, pg2, pg3);
                pact = act; psym = sym; prank = rank;
                pg0 = g0; pg1 = g1; pg2 = g2; pg3 = g3;
            }
        }
        {   // drain the pipeline: the last step's update
            double2 *prow = reinterpret_cast<double2 *>(cntw + psym * 4);
            if (pact && prank == 0) {
                double2 r01 = prow[0], r23 = prow[1];
                r01.x += pg0; r01.y += pg1; r23.x += pg2; r23.y += pg3;
                prow[0] = r01; prow[1] = r23;
            }
            __syncwarp();
            cnt_update4_rest(cntw, pact, psym, prank, pg0, pg1, pg2, pg3);
        }
        if (st.imprecise) {  // sticky hand-over; the host redoes this E-step once (hmmb_bw_iterate)
            flag[bk.first + lane] = 1;
            atomicAdd(new_flags, 1);
        }
    }

    // ---- CTA flush: deterministic (fixed-order) reduction into this CTA's partial
    double Xf[16];
    if (BIDIAG) {
#pragma unroll
        for (int q = 0; q < 16; ++q) Xf[q] = 0.0;
        Xf[0] = st.X[0]; Xf[5] = st.X[1]; Xf[10] = st.X[2]; Xf[15] = st.X[3];
        Xf[1] = st.X[4]; Xf[6] = st.X[5]; Xf[11] = st.X[6];
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) Xf[q] = st.X[q];
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        double v = Xf[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[warp][4 + q] = v;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double v = mypi[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[warp][q] = v;
    }
    const unsigned seenW = __reduce_or_sync(0xffffffffu, st.seenX);
    if (lane == 0) atomicOr(&sSeen, seenW);
    __syncthreads();
    if (tid < 20) {
        double v = ((sRed[0][tid] + sRed[1][tid]) + sRed[2][tid]) + sRed[3][tid];
        if (tid >= 4) {
            const int q = tid - 4;
            const double aij = __ldg(A + (size_t)cw.word * 16 + q);
            double val = aij > 0.0 ? aij * v : 0.0;  // impossible transitions: ignore whatever piled up
            if (val == 0.0 && aij > 0.0 && ((sSeen >> q) & 1u)) val = tiny_pos();
            v = val;
        }
        part[tid] = v;
    }
    for (int e = tid; e < M * 4; e += BW_THREADS)
        part[20 + e] = ((sCnt[e] + sCnt[(size_t)M * 4 + e]) + sCnt[(size_t)M * 8 + e]) + sCnt[(size_t)M * 12 + e];
}

}  // namespace hmmb
