// Baum-Welch kernels (E-step forward / backward+accumulate, reduce, M-step, finalize).
//
// Replaces HMM/hmm_training.py:265-541 of the reference.  Two kernel families:
//   * N == 4 (the reference's hard-coded state count, hmm_training.py:226): one sequence per
//     thread, 32 sequences per warp in lock-step, A in registers, B^T of the CTA's word in
//     shared memory, alpha-hat spilled to HBM as fully coalesced 16-byte stores
//     [block][t][half][lane], backward pass fused with the gamma / xi / emission-count
//     accumulation (beta never leaves registers), emission counts in warp-private shared
//     memory copies (no atomics: lanes that hit the same codeword take turns, found with
//     MATCH.ANY), one deterministic partial per CTA.
//   * generic N <= 32: NP = 4/8/16/32 lanes per sequence (lane = state), alpha/v broadcast
//     through a per-warp shared staging buffer, accumulators via fp64 RED to L2.
#pragma once

#include "hmm_device.cuh"

namespace hmmb {

constexpr int BW_THREADS = 128;  // 4 warps per CTA in every E-step kernel
constexpr int BW_WARPS = BW_THREADS / 32;

// ---------------------------------------------------------------- repack (N = 4 path)
// Gathers the ragged codeword stream into the blocked layout the N = 4 kernels read:
// obs_blk[blk.obs_base + chunk*32 + lane] = 16 bytes = SPC consecutive symbols of lane's
// sequence.  Also checks codeword < M (the reference would raise IndexError).
template <typename InT, typename SymT>
__global__ void k_repack_blocks(const InT *__restrict__ obs, const int64_t *__restrict__ off_sorted,
                                const int32_t *__restrict__ len_sorted, const Blk *__restrict__ blks, int nblk,
                                uint4 *__restrict__ obs_blk, int M, int *__restrict__ bad) {
    constexpr int SPC = Sym<SymT>::SPC;
    const int b = blockIdx.x;
    if (b >= nblk) return;
    const Blk bk = blks[b];
    const int nch = (bk.tmax + SPC - 1) / SPC;
    for (int e = threadIdx.x; e < nch * 32; e += blockDim.x) {
        const int c = e >> 5, lane = e & 31;
        SymT vals[SPC];
#pragma unroll
        for (int s = 0; s < SPC; ++s) vals[s] = 0;
        if (lane < bk.nseq) {
            const int T = len_sorted[bk.first + lane];
            const InT *src = obs + off_sorted[bk.first + lane];
#pragma unroll
            for (int s = 0; s < SPC; ++s) {
                const int t = c * SPC + s;
                if (t < T) {
                    unsigned long long v = (unsigned long long)src[t];
                    if (v >= (unsigned long long)M) { atomicOr(bad, 1); v = 0; }
                    vals[s] = (SymT)v;
                }
            }
        }
        uint4 w;
        if constexpr (SPC == 16) {
            unsigned r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                r[q] = (unsigned)vals[4 * q] | ((unsigned)vals[4 * q + 1] << 8) | ((unsigned)vals[4 * q + 2] << 16) |
                       ((unsigned)vals[4 * q + 3] << 24);
            w = make_uint4(r[0], r[1], r[2], r[3]);
        } else {
            unsigned r[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) r[q] = (unsigned)vals[2 * q] | ((unsigned)vals[2 * q + 1] << 16);
            w = make_uint4(r[0], r[1], r[2], r[3]);
        }
        obs_blk[bk.obs_base + e] = w;
    }
}

// generic path: convert to the canonical symbol width, validate range
template <typename InT, typename SymT>
__global__ void k_convert_obs(const InT *__restrict__ obs, int64_t n, SymT *__restrict__ out, int M,
                              int *__restrict__ bad) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long v = (unsigned long long)obs[i];
        if (v >= (unsigned long long)M) { atomicOr(bad, 1); v = 0; }
        out[i] = (SymT)v;
    }
}

// ---------------------------------------------------------------- N = 4 forward
template <typename SymT>
__global__ void __launch_bounds__(BW_THREADS)
k_bw_fwd4(const CtaWork *__restrict__ work, const Blk *__restrict__ blks, const uint4 *__restrict__ obs_blk,
          const int32_t *__restrict__ len_sorted, const double *__restrict__ pi, const double *__restrict__ A,
          const double *__restrict__ Bt, int M, double2 *__restrict__ spill, double *__restrict__ ll_seq,
          const int32_t *__restrict__ active, uint8_t *__restrict__ flag) {
    extern __shared__ double sB[];
    const CtaWork cw = work[blockIdx.x];
    if (!active[cw.word]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const double2 *src = reinterpret_cast<const double2 *>(Bt + (size_t)cw.word * M * 4);
        double2 *dst = reinterpret_cast<double2 *>(sB);
        for (int e = tid; e < M * 2; e += BW_THREADS) dst[e] = __ldg(src + e);
    }
    double a[16], p[4];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = __ldg(A + (size_t)cw.word * 16 + q);
#pragma unroll
    for (int q = 0; q < 4; ++q) p[q] = __ldg(pi + (size_t)cw.word * 4 + q);
    __syncthreads();
    for (int b = cw.blk_begin + warp; b < cw.blk_end; b += BW_WARPS) {
        const Blk bk = blks[b];
        int T = lane < bk.nseq ? len_sorted[bk.first + lane] : 0;
        if (T > 0 && flag[bk.first + lane]) T = 0;  // handled by the exact log-space kernel
        const double ll = fwd4_run<SymT, true>(T, bk.tmax, obs_blk + bk.obs_base + lane, sB, a, p,
                                               spill + bk.spill_base * 64 + lane);
        if (T > 0) {
            ll_seq[bk.first + lane] = ll;
            if (ll != ll) flag[bk.first + lane] = 1;  // precision guard: hand over (sticky)
        }
    }
}

// ---------------------------------------------------------------- N = 4 backward + accumulate
// Warp-private emission-count update: every active lane adds its 4 gammas to row `sym` of
// the warp's count copy.  Lanes that share a codeword in this step take turns in lane order
// (deterministic), found with MATCH.ANY; no atomics.
__device__ __forceinline__ void cnt_update4(double *__restrict__ cw, bool act, unsigned sym, int lane, double g0,
                                            double g1, double g2, double g3) {
    const unsigned key = act ? sym : (0x10000u | (unsigned)lane);
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const int maxrank = __reduce_max_sync(0xffffffffu, rank);
    double2 *row = reinterpret_cast<double2 *>(cw + (act ? sym : 0u) * 4);
    for (int r = 0; r <= maxrank; ++r) {
        if (act && rank == r) {
            double2 c01 = row[0], c23 = row[1];
            c01.x += g0; c01.y += g1; c23.x += g2; c23.y += g3;
            row[0] = c01; row[1] = c23;
        }
        __syncwarp();
    }
}

// Partial layout per CTA (and accumulator layout per word): [pi N][xi N*N][cnt M*N].
template <typename SymT>
__global__ void __launch_bounds__(BW_THREADS, 3)
k_bw_bwd4(const CtaWork *__restrict__ work, const Blk *__restrict__ blks, const uint4 *__restrict__ obs_blk,
          const int32_t *__restrict__ len_sorted, const double *__restrict__ A, const double *__restrict__ Bt, int M,
          const double2 *__restrict__ spill, const double *__restrict__ ll_seq, const int32_t *__restrict__ active,
          double *__restrict__ partials, int64_t pstride, uint8_t *__restrict__ flag, int32_t *__restrict__ new_flags) {
    constexpr int SPC = Sym<SymT>::SPC;
    extern __shared__ double smem[];
    double *sB = smem;                        // [M][4]
    double *sCnt = smem + (size_t)M * 4;      // [4 warps][M][4]
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
        for (int e = tid; e < M * 4 * BW_WARPS; e += BW_THREADS) sCnt[e] = 0.0;
        if (tid == 0) sSeen = 0u;
    }
    double a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = __ldg(A + (size_t)cw.word * 16 + q);
    __syncthreads();

    double *cntw = sCnt + (size_t)warp * M * 4;
    double X[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) X[q] = 0.0;
    double pa0 = 0.0, pa1 = 0.0, pa2 = 0.0, pa3 = 0.0;
    unsigned seenX = 0u;

    for (int b = cw.blk_begin + warp; b < cw.blk_end; b += BW_WARPS) {
        const Blk bk = blks[b];
        int T = 0;
        if (lane < bk.nseq) {
            T = len_sorted[bk.first + lane];
            if (!(ll_seq[bk.first + lane] > neg_inf())) T = 0;  // impossible sequence: contributes nothing (:391-394)
            if (flag[bk.first + lane]) T = 0;                   // exact log-space kernel did this one
        }
        bool imprecise = false;
        const uint4 *op = obs_blk + bk.obs_base + lane;
        const double2 *sp = spill + bk.spill_base * 64 + lane;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;  // v_j = b_j(o_{t+1}) * beta-hat_{t+1}(j)
        double2 nx01 = make_double2(0.0, 0.0), nx23 = make_double2(0.0, 0.0);
        if (T > 0) {
            nx01 = __ldcs(sp + (size_t)(T - 1) * 64);
            nx23 = __ldcs(sp + (size_t)(T - 1) * 64 + 32);
        }
        const int nch = (bk.tmax + SPC - 1) / SPC;
        for (int c = nch - 1; c >= 0; --c) {
            uint4 w = __ldg(op + (size_t)c * 32);
#pragma unroll 2
            for (int s = SPC - 1; s >= 0; --s) {
                const int t = c * SPC + s;
                const unsigned sym = Sym<SymT>::pop_back(w);
                if (t >= bk.tmax) continue;  // warp-uniform
                const bool act = t < T;
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                if (act) {
                    const double al0 = nx01.x, al1 = nx01.y, al2 = nx23.x, al3 = nx23.y;
                    if (t > 0) {
                        nx01 = __ldcs(sp + (size_t)(t - 1) * 64);
                        nx23 = __ldcs(sp + (size_t)(t - 1) * 64 + 32);
                    }
                    const bool last = (t == T - 1);
                    // h = beta-hat_t (power-of-two scaled so that sum_i h_i is in [1,2));
                    // w_j = v_j under the same scale, so xi_t = al_i a_ij w_j / sum_i al_i h_i (:397-410)
                    double h0 = 1.0, h1 = 1.0, h2 = 1.0, h3 = 1.0;  // log beta_{T-1} = 0 (:363)
                    double w0 = 0.0, w1 = 0.0, w2 = 0.0, w3 = 0.0;
                    if (!last) {
                        // un-normalised beta_t(i) = sum_j a_ij b_j(o_{t+1}) beta_{t+1}(j)  (:163-199)
                        double q0 = a[0] * v0 + a[1] * v1 + a[2] * v2 + a[3] * v3;
                        double q1 = a[4] * v0 + a[5] * v1 + a[6] * v2 + a[7] * v3;
                        double q2 = a[8] * v0 + a[9] * v1 + a[10] * v2 + a[11] * v3;
                        double q3 = a[12] * v0 + a[13] * v1 + a[14] * v2 + a[15] * v3;
                        if (maybe_zero(q0) | maybe_zero(q1) | maybe_zero(q2) | maybe_zero(q3)) {
#define HMMB_FIX_Q(I, Q)                                                                            \
    if (Q == 0.0 && ((a[4 * I] > 0.0 && v0 > 0.0) || (a[4 * I + 1] > 0.0 && v1 > 0.0) ||             \
                     (a[4 * I + 2] > 0.0 && v2 > 0.0) || (a[4 * I + 3] > 0.0 && v3 > 0.0)))          \
        Q = tiny_pos();
                            HMMB_FIX_Q(0, q0) HMMB_FIX_Q(1, q1) HMMB_FIX_Q(2, q2) HMMB_FIX_Q(3, q3)
#undef HMMB_FIX_Q
                        }
                        const double qs = (q0 + q1) + (q2 + q3);
                        const double sc = qs > 0.0 ? pow2_rescale_noacc(qs) : 1.0;
                        h0 = q0 * sc; h1 = q1 * sc; h2 = q2 * sc; h3 = q3 * sc;
                        w0 = v0 * sc; w1 = v1 * sc; w2 = v2 * sc; w3 = v3 * sc;
                    }
                    // gamma_t(i) = alpha_t(i) beta_t(i) / sum_i alpha_t(i) beta_t(i)   (:389-394)
                    g0 = al0 * h0; g1 = al1 * h1; g2 = al2 * h2; g3 = al3 * h3;
                    double norm = (g0 + g1) + (g2 + g3);
                    double r = 1.0 / norm;
                    double u0 = al0, u1 = al1, u2 = al2, u3 = al3;
                    const bool slow = maybe_zero(g0) | maybe_zero(g1) | maybe_zero(g2) | maybe_zero(g3) | !(norm >= TINY_STEP);
                    if (slow) {
                        if (!(norm >= TINY_STEP)) {
                            // forward and backward mass sit on (almost) disjoint states: redo the
                            // products 2^1000 larger so they neither underflow nor blow up 1/norm;
                            // the bits may already be gone, so also hand the sequence over
                            imprecise = true;
                            const double big = 0x1p500;
                            u0 = al0 * big; u1 = al1 * big; u2 = al2 * big; u3 = al3 * big;
                            g0 = u0 * (h0 * big); g1 = u1 * (h1 * big); g2 = u2 * (h2 * big); g3 = u3 * (h3 * big);
                            w0 *= big; w1 *= big; w2 *= big; w3 *= big;
                            norm = (g0 + g1) + (g2 + g3);
                            r = norm > 0.0 ? 1.0 / norm : 0.0;
                        }
                        g0 *= r; g1 *= r; g2 *= r; g3 *= r;
                        if (g0 == 0.0 && al0 > 0.0 && h0 > 0.0) g0 = tiny_pos();
                        if (g1 == 0.0 && al1 > 0.0 && h1 > 0.0) g1 = tiny_pos();
                        if (g2 == 0.0 && al2 > 0.0 && h2 > 0.0) g2 = tiny_pos();
                        if (g3 == 0.0 && al3 > 0.0 && h3 > 0.0) g3 = tiny_pos();
                    } else {
                        g0 *= r; g1 *= r; g2 *= r; g3 *= r;
                    }
                    if (!last) {
                        // a_ij is factored out of the time sum and applied once at the flush
                        u0 *= r; u1 *= r; u2 *= r; u3 *= r;
                        X[0] = fma(u0, w0, X[0]);  X[1] = fma(u0, w1, X[1]);  X[2] = fma(u0, w2, X[2]);  X[3] = fma(u0, w3, X[3]);
                        X[4] = fma(u1, w0, X[4]);  X[5] = fma(u1, w1, X[5]);  X[6] = fma(u1, w2, X[6]);  X[7] = fma(u1, w3, X[7]);
                        X[8] = fma(u2, w0, X[8]);  X[9] = fma(u2, w1, X[9]);  X[10] = fma(u2, w2, X[10]); X[11] = fma(u2, w3, X[11]);
                        X[12] = fma(u3, w0, X[12]); X[13] = fma(u3, w1, X[13]); X[14] = fma(u3, w2, X[14]); X[15] = fma(u3, w3, X[15]);
                        if (slow | maybe_zero(al0) | maybe_zero(al1) | maybe_zero(al2) | maybe_zero(al3) | maybe_zero(v0) |
                            maybe_zero(v1) | maybe_zero(v2) | maybe_zero(v3)) {
                            const unsigned mv = (v0 > 0.0 ? 1u : 0u) | (v1 > 0.0 ? 2u : 0u) | (v2 > 0.0 ? 4u : 0u) | (v3 > 0.0 ? 8u : 0u);
                            seenX |= (al0 > 0.0 ? mv : 0u) | (al1 > 0.0 ? mv << 4 : 0u) | (al2 > 0.0 ? mv << 8 : 0u) |
                                     (al3 > 0.0 ? mv << 12 : 0u);
                        } else {
                            seenX = 0xffffu;
                        }
                    }
                    if (t == 0) { pa0 += g0; pa1 += g1; pa2 += g2; pa3 += g3; }  // (:415-426)
                    // v_j = b_j(o_t) beta-hat_t(j) for step t-1
                    const double2 b01 = *reinterpret_cast<const double2 *>(sB + sym * 4);
                    const double2 b23 = *reinterpret_cast<const double2 *>(sB + sym * 4 + 2);
                    v0 = b01.x * h0; v1 = b01.y * h1; v2 = b23.x * h2; v3 = b23.y * h3;
                    const double vs = (v0 + v1) + (v2 + v3);
                    if (!(vs >= TINY_STEP) | maybe_zero(v0) | maybe_zero(v1) | maybe_zero(v2) | maybe_zero(v3)) {
                        if (!(vs >= TINY_STEP)) {
                            // tiny emission column: exponent-split products (the scale of v is free)
                            double o[4];
                            int E;
                            if (exact_products4(h0, h1, h2, h3, b01.x, b01.y, b23.x, b23.y, o, &E) == 2) imprecise = true;
                            v0 = o[0]; v1 = o[1]; v2 = o[2]; v3 = o[3];
                        } else {
                            if (v0 == 0.0 && b01.x > 0.0 && h0 > 0.0) v0 = tiny_pos();
                            if (v1 == 0.0 && b01.y > 0.0 && h1 > 0.0) v1 = tiny_pos();
                            if (v2 == 0.0 && b23.x > 0.0 && h2 > 0.0) v2 = tiny_pos();
                            if (v3 == 0.0 && b23.y > 0.0 && h3 > 0.0) v3 = tiny_pos();
                        }
                    }
                }
                cnt_update4(cntw, act, sym, lane, g0, g1, g2, g3);  // (:460-500 numerators)
            }
        }
        if (imprecise) {  // sticky hand-over; the host redoes this E-step once (hmmb_bw_iterate)
            flag[bk.first + lane] = 1;
            atomicAdd(new_flags, 1);
        }
    }

    // ---- CTA flush: deterministic (fixed-order) reduction into this CTA's partial
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        double v = X[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sRed[warp][4 + q] = v;
    }
    {
        double pv[4] = {pa0, pa1, pa2, pa3};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            double v = pv[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) sRed[warp][q] = v;
        }
    }
    const unsigned seenW = __reduce_or_sync(0xffffffffu, seenX);
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

// ---------------------------------------------------------------- generic-N forward
// Each group of NP lanes walks the contiguous range of sequences
// [group * per_group, (group+1) * per_group) of the word-sorted order.
template <int NP, typename SymT>
__global__ void __launch_bounds__(BW_THREADS)
k_bw_fwdG(const SymT *__restrict__ obs, const int64_t *__restrict__ off_sorted, const int32_t *__restrict__ len_sorted,
          const int32_t *__restrict__ word_sorted, const int64_t *__restrict__ foff_sorted, int64_t R,
          int64_t per_group, int N, int M, const double *__restrict__ pi, const double *__restrict__ A,
          const double *__restrict__ Bt, double *__restrict__ spill, double *__restrict__ ll_seq,
          const int32_t *__restrict__ active, uint8_t *__restrict__ flag) {
    __shared__ double sStage[BW_WARPS][128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int GPW = 32 / NP;
    const int j = lane % NP, gbase = lane - j;
    const int64_t group = ((int64_t)blockIdx.x * BW_WARPS + warp) * GPW + lane / NP;
    double acol[NP];
    double pj = 0.0;
    int curw = -1;
    for (int64_t k = 0; k < per_group; ++k) {
        const int64_t r = group * per_group + k;
        int T = 0, w = 0;
        if (r < R) {
            w = word_sorted[r];
            if (active[w] && !flag[r]) T = len_sorted[r];
        }
        const int Tw = warp_max_int(T);
        if (Tw == 0) continue;  // warp-uniform
        if (T > 0 && w != curw) {
            curw = w;
#pragma unroll
            for (int i = 0; i < NP; ++i) acol[i] = (i < N && j < N) ? __ldg(A + ((size_t)w * N + i) * N + j) : 0.0;
            pj = j < N ? __ldg(pi + (size_t)w * N + j) : 0.0;
        }
        const SymT *o = obs + (T > 0 ? off_sorted[r] : 0);
        double *sp = spill + (T > 0 ? foff_sorted[r] * N : 0);
        const double ll = fwdG_run<NP, SymT, true>(T, Tw, N, j, gbase, lane, o, Bt + (size_t)w * M * N, acol, pj, sp,
                                                   sStage[warp]);
        if (T > 0 && j == 0) {
            ll_seq[r] = ll;
            if (ll != ll) flag[r] = 1;  // precision guard: hand over to the exact kernel (sticky)
        }
    }
}

// ---------------------------------------------------------------- generic-N backward + accumulate
// Lane i = state i of its group's sequence; holds row i of A and row i of the xi accumulator.
// Accumulator layout per word: [pi N][xi N*N][cnt M*N] (fp64 RED into `accum`).
template <int NP, typename SymT>
__global__ void __launch_bounds__(BW_THREADS)
k_bw_bwdG(const SymT *__restrict__ obs, const int64_t *__restrict__ off_sorted, const int32_t *__restrict__ len_sorted,
          const int32_t *__restrict__ word_sorted, const int64_t *__restrict__ foff_sorted, int64_t R,
          int64_t per_group, int N, int M, const double *__restrict__ A, const double *__restrict__ Bt,
          const double *__restrict__ spill, const double *__restrict__ ll_seq, const int32_t *__restrict__ active,
          double *__restrict__ accum, int64_t astride, uint8_t *__restrict__ flag, int32_t *__restrict__ new_flags) {
    __shared__ double sStage[BW_WARPS][128];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int GPW = 32 / NP;
    const int i = lane % NP, gbase = lane - i;
    const unsigned gmask = (NP == 32) ? 0xffffffffu : ((1u << NP) - 1u);
    const int64_t group = ((int64_t)blockIdx.x * BW_WARPS + warp) * GPW + lane / NP;
    double arow[NP], Xrow[NP];
    int curw = -1;
    double *stage = sStage[warp];
    for (int64_t k = 0; k < per_group; ++k) {
        const int64_t r = group * per_group + k;
        int T = 0, w = 0;
        if (r < R) {
            w = word_sorted[r];
            if (active[w] && !flag[r] && ll_seq[r] > neg_inf()) T = len_sorted[r];
        }
        const int Tw = warp_max_int(T);
        if (Tw == 0) continue;
        bool imprecise = false;
        if (T > 0 && w != curw) {
            curw = w;
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) arow[jj] = (i < N && jj < N) ? __ldg(A + ((size_t)w * N + i) * N + jj) : 0.0;
        }
#pragma unroll
        for (int jj = 0; jj < NP; ++jj) Xrow[jj] = 0.0;
        unsigned seenrow = 0u;
        const SymT *o = obs + (T > 0 ? off_sorted[r] : 0);
        const double *sp = spill + (T > 0 ? foff_sorted[r] * N : 0);
        const double *Btw = Bt + (size_t)w * M * N;
        double *acc = accum + (size_t)w * astride;
        double v = 0.0;
        int par = 0;
        stage[par * 32 + lane] = 0.0;
        __syncwarp();
        const unsigned gm = gmask << gbase;
        for (int t = Tw - 1; t >= 0; --t) {
            const bool act = t < T;
            const bool last = (t == T - 1);
            const unsigned vmask = (__ballot_sync(0xffffffffu, v > 0.0) >> gbase) & gmask;
            double q = 0.0;
            if (act) {
                if (last) {
                    q = i < N ? 1.0 : 0.0;  // log beta_{T-1} = 0 (:363)
                } else {
                    const double *sv = stage + par * 32 + gbase;
#pragma unroll
                    for (int jj = 0; jj < NP; ++jj) q = fma(arow[jj], sv[jj], q);
                    if (q == 0.0 && i < N) {
                        bool reach = false;
                        for (int jj = 0; jj < NP; ++jj) reach |= (arow[jj] > 0.0 && sv[jj] > 0.0);
                        if (reach) q = tiny_pos();
                    }
                }
            }
            const double qs = group_sum<NP>(q);
            double h = 0.0, al = 0.0, g = 0.0, sc = 1.0;
            if (act) {
                if (!last && qs > 0.0) sc = pow2_rescale_noacc(qs);
                h = q * sc;  // beta-hat_t(i), group sum in [1,2)
                al = i < N ? sp[(size_t)t * N + i] : 0.0;
                g = al * h;
            }
            double norm = group_sum<NP>(g);
            unsigned sym = 0;
            if (act) {
                double u = al, wscale = 1.0;
                if (!(norm >= TINY_STEP)) {  // group-uniform: see the N = 4 kernel
                    imprecise = true;
                    const double big = 0x1p500;
                    u = al * big;
                    g = u * (h * big);
                    wscale = big;
                    norm = g;
#pragma unroll
                    for (int o = NP / 2; o > 0; o >>= 1) norm += __shfl_xor_sync(gm, norm, o);
                }
                const double r1 = norm > 0.0 ? 1.0 / norm : 0.0;
                g *= r1;  // gamma_t(i) (:389-394)
                if (g == 0.0 && al > 0.0 && h > 0.0) g = tiny_pos();
                sym = o[t];
                if (!last) {
                    u *= r1;
                    const double *sv = stage + par * 32 + gbase;
#pragma unroll
                    for (int jj = 0; jj < NP; ++jj) Xrow[jj] = fma(u, (sv[jj] * sc) * wscale, Xrow[jj]);
                    if (al > 0.0) seenrow |= vmask;
                }
                if (i < N && g > 0.0) {
                    atomicAdd(acc + N + N * N + (size_t)sym * N + i, g);  // (:460-500 numerators)
                    if (t == 0) atomicAdd(acc + i, g);                   // (:415-426)
                }
            }
            // v_i = b_i(o_t) beta-hat_t(i) for step t-1
            double b = 0.0;
            if (act) {
                b = i < N ? __ldg(Btw + (size_t)sym * N + i) : 0.0;
                v = b * h;
                if (v == 0.0 && b > 0.0 && h > 0.0) v = tiny_pos();
            }
            const unsigned okm = (__ballot_sync(0xffffffffu, v >= TINY_STEP) >> gbase) & gmask;
            if (act && okm == 0u) {
                // every v of the group is tiny: exponent-split products (the scale of v is free)
                double m = 0.0;
                int e = INT_MIN;
                if (b > 0.0 && h > 0.0) frexp_prod(b, h, m, e);
                int E = e;
#pragma unroll
                for (int o2 = NP / 2; o2 > 0; o2 >>= 1) E = max(E, __shfl_xor_sync(gm, E, o2));
                v = 0.0;
                if (e != INT_MIN) {
                    v = ldexp(m, max(e - E, -1200));
                    if (v == 0.0) v = tiny_pos();
                    if (e >= E - 64 && h < SUBNORMAL_LIMIT) imprecise = true;
                }
            }
            par ^= 1;
            stage[par * 32 + lane] = v;
            __syncwarp();
        }
        if (T > 0 && i < N) {
#pragma unroll
            for (int jj = 0; jj < NP; ++jj) {
                if (jj < N) {
                    double val = arow[jj] * Xrow[jj];
                    if (val == 0.0 && arow[jj] > 0.0 && ((seenrow >> jj) & 1u)) val = tiny_pos();
                    if (arow[jj] > 0.0 && val > 0.0) atomicAdd(acc + N + (size_t)i * N + jj, val);
                }
            }
        }
        {
            const unsigned bal = __ballot_sync(0xffffffffu, imprecise && T > 0);
            if (((bal >> gbase) & gmask) != 0u && T > 0 && i == 0) {
                flag[r] = 1;
                atomicAdd(new_flags, 1);
            }
            imprecise = false;
        }
        v = 0.0;
    }
}

// ---------------------------------------------------------------- exact log-space E-step (rare)
// One warp per sequence handed over by the precision guard (flag[r] != 0): the reference's
// log-space forward / backward, contributions added to `accum` with fp64 atomics.  Runs after
// the fast forward kernel and before the reduce kernel.  base_sorted[r]: codeword offset of the
// sequence (generic: into the linear stream; N = 4 path: uint4 row of its lane in obs_blk).
template <typename SymT, bool BLOCKED>
__global__ void __launch_bounds__(BW_THREADS)
k_bw_exact(const void *__restrict__ obs, const int64_t *__restrict__ base_sorted, const int32_t *__restrict__ len_sorted,
           const int32_t *__restrict__ word_sorted, int64_t R, int N, int M, const double *__restrict__ pi,
           const double *__restrict__ A, const double *__restrict__ Bt, double *__restrict__ ll_seq,
           const int32_t *__restrict__ active, const uint8_t *__restrict__ flag, double *__restrict__ scratch,
           int64_t scratch_stride, double *__restrict__ accum, int64_t astride, int64_t *__restrict__ n_exact) {
    const int lane = threadIdx.x & 31;
    const int64_t gw = (int64_t)blockIdx.x * BW_WARPS + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * BW_WARPS;
    double *sc = scratch + gw * scratch_stride;
    for (int64_t base = gw * 32; base < R; base += nw * 32) {
        const int64_t r = base + lane;
        const bool f = r < R && flag[r] && active[word_sorted[r]];
        unsigned m = __ballot_sync(0xffffffffu, f);
        while (m) {
            const int l = __ffs(m) - 1;
            m &= m - 1;
            const int64_t rr = base + l;
            const int w = word_sorted[rr], T = len_sorted[rr];
            const double *piw = pi + (size_t)w * N, *Aw = A + (size_t)w * N * N, *Btw = Bt + (size_t)w * M * N;
            double logP;
            if (BLOCKED) {
                BlkObs<SymT> o{reinterpret_cast<const uint4 *>(obs) + base_sorted[rr]};
                logP = exact_forward(T, N, lane, o, piw, Aw, Btw, sc);
                __syncwarp();
                if (logP > neg_inf()) exact_backward_accumulate(T, N, lane, o, Aw, Btw, sc, logP, accum + (size_t)w * astride);
            } else {
                LinObs<SymT> o{reinterpret_cast<const SymT *>(obs) + base_sorted[rr]};
                logP = exact_forward(T, N, lane, o, piw, Aw, Btw, sc);
                __syncwarp();
                if (logP > neg_inf()) exact_backward_accumulate(T, N, lane, o, Aw, Btw, sc, logP, accum + (size_t)w * astride);
            }
            if (lane == 0) {
                ll_seq[rr] = logP;
                atomicAdd(reinterpret_cast<unsigned long long *>(n_exact), 1ULL);
            }
            __syncwarp();
        }
    }
}

// ---------------------------------------------------------------- reduce + LL statistics
// One CTA per word.  (N = 4 path) sums the word's CTA partials in fixed order into accum;
// (both) writes the word's sequence count and this rank's (max, sum exp) of the
// per-sequence log-likelihoods, the two halves of log_sum_exp (:503), into
// llstats[rank][w][2].
constexpr int RED_THREADS = 256;

__global__ void __launch_bounds__(RED_THREADS)
k_bw_reduce(const double *__restrict__ partials, int64_t pstride, const int32_t *__restrict__ cta_begin,
            const double *__restrict__ ll_seq, const int64_t *__restrict__ seq_begin, double *__restrict__ accum,
            int64_t astride, int64_t nacc, double *__restrict__ llstats, int rank, int W,
            const int32_t *__restrict__ active) {
    __shared__ double sM[RED_THREADS / 32], sS[RED_THREADS / 32];
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!active[w]) return;
    double *acc = accum + (size_t)w * astride;
    if (partials) {
        const int c0 = cta_begin[w], c1 = cta_begin[w + 1];
        for (int64_t e = tid; e < nacc; e += RED_THREADS) {
            double s = 0.0;
            for (int c = c0; c < c1; ++c) s += partials[(size_t)c * pstride + e];
            acc[e] += s;  // the exact log-space kernel may already have added flagged sequences
        }
    }
    const int64_t r0 = seq_begin[w], r1 = seq_begin[w + 1];
    double m = neg_inf();
    for (int64_t r = r0 + tid; r < r1; r += RED_THREADS) m = fmax(m, ll_seq[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) sM[warp] = m;
    __syncthreads();
    m = sM[0];
    for (int q = 1; q < RED_THREADS / 32; ++q) m = fmax(m, sM[q]);
    double s = 0.0;
    if (m > neg_inf())
        for (int64_t r = r0 + tid; r < r1; r += RED_THREADS) {
            const double l = ll_seq[r];
            if (l > neg_inf()) s += exp(l - m);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sS[warp] = s;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int q = 0; q < RED_THREADS / 32; ++q) tot += sS[q];
        acc[nacc] = (double)(r1 - r0);  // sequences of this word on this rank (num_recordings, :268)
        llstats[((size_t)rank * W + w) * 2 + 0] = m;
        llstats[((size_t)rank * W + w) * 2 + 1] = tot;
    }
}

// ---------------------------------------------------------------- M-step
// One CTA per word; replaces hmm_training.py:412-514.
__global__ void __launch_bounds__(RED_THREADS)
k_bw_mstep(const double *__restrict__ accum, int64_t astride, const double *__restrict__ llstats, int world, int W,
           int N, int M, double *__restrict__ pi, double *__restrict__ A, double *__restrict__ Bt,
           int32_t *__restrict__ active, int32_t *__restrict__ iters, double *__restrict__ prev_ll,
           double *__restrict__ ll_hist, int hist_cap, double eps, int max_iter, int32_t *__restrict__ any_active) {
    __shared__ double sDen[HMMB_MAX_STATES];
    __shared__ double sPart[RED_THREADS / 32][HMMB_MAX_STATES];
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (!active[w]) return;
    const double *acc = accum + (size_t)w * astride;
    const double *xi = acc + N;
    const double *cnt = acc + N + N * N;
    const double nseq = acc[N + N * N + (size_t)M * N];

    // B denominators: sum over all t of gamma_t(j) = column sums of the counts (:462-472)
    for (int j = 0; j < N; ++j) {
        double s = 0.0;
        for (int k = tid; k < M; k += RED_THREADS) s += cnt[(size_t)k * N + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sPart[warp][j] = s;
    }
    __syncthreads();
    if (tid < N) {
        double s = 0.0;
        for (int q = 0; q < RED_THREADS / 32; ++q) s += sPart[q][tid];
        sDen[tid] = s;
    }
    __syncthreads();
    // B (:474-497): no finite term -> 1e-20 floor; empty denominator -> row stays -inf (0)
    for (int e = tid; e < M * N; e += RED_THREADS) {
        const int j = e % N;
        const double den = sDen[j];
        const double c = cnt[e];
        double b = 0.0;
        if (den > 0.0) {
            if (c > 0.0) {
                b = c / den;
                if (b == 0.0) b = tiny_pos();
            } else {
                b = 1e-20;
            }
        }
        Bt[(size_t)w * M * N + e] = b;
    }
    // A (:429-457): a_ij = sum xi / sum_{t<T-1} gamma(i); the denominator equals the row sum of xi
    if (tid < N) {
        const int i = tid;
        double den = 0.0;
        for (int j = 0; j < N; ++j) den += xi[i * N + j];
        for (int j = 0; j < N; ++j) {
            double x = xi[i * N + j];
            double v = 0.0;
            if (den > 0.0 && x > 0.0) {
                v = x / den;
                if (v == 0.0) v = tiny_pos();
            }
            A[((size_t)w * N + i) * N + j] = v;
        }
        // pi (:415-426): sum_r gamma_0^r(i) / num_recordings
        const double p = acc[i];
        double pv = 0.0;
        if (p > 0.0) {
            pv = p / nseq;
            if (pv == 0.0) pv = tiny_pos();
        }
        pi[(size_t)w * N + i] = pv;
    }
    if (tid == 0) {
        // convergence statistic: log_sum_exp over all sequences of the word (:503-508)
        double mx = neg_inf();
        for (int r = 0; r < world; ++r) mx = fmax(mx, llstats[((size_t)r * W + w) * 2]);
        double cur = neg_inf();
        if (mx > neg_inf()) {
            double s = 0.0;
            for (int r = 0; r < world; ++r) {
                const double m = llstats[((size_t)r * W + w) * 2];
                if (m > neg_inf()) s += llstats[((size_t)r * W + w) * 2 + 1] * exp(m - mx);
            }
            cur = mx + log(s);
        }
        const double prev = prev_ll[w];
        const double diff = (prev > neg_inf()) ? fabs(cur - prev) : pos_inf();
        int it = iters[w];
        if (it < hist_cap) ll_hist[(size_t)w * hist_cap + it] = cur;
        ++it;
        iters[w] = it;
        prev_ll[w] = cur;
        const int act = (diff >= eps && it < max_iter) ? 1 : 0;  // loop condition (:346)
        active[w] = act;
        if (act) atomicOr(any_active, 1);
    }
}

// ---------------------------------------------------------------- exit normalisation (:524-539)
// finalize: safe_exp of the log parameters (:524-526) then the row normalisations.  A value
// held at the smallest denormal is our marker for "finite log value below the double range"
// (the reference's exp underflows to exactly 0.0 there), so it is flushed to 0 on exit.
__device__ __forceinline__ double exit_value(double v, int finalize) {
    return (finalize && v == tiny_pos()) ? 0.0 : v;
}

__global__ void __launch_bounds__(RED_THREADS)
k_bw_finalize(const double *__restrict__ pi, const double *__restrict__ A, const double *__restrict__ Bt, int N, int M,
              int finalize, double *__restrict__ pi_out, double *__restrict__ A_out, double *__restrict__ B_out) {
    __shared__ double sDen[HMMB_MAX_STATES];
    __shared__ double sPart[RED_THREADS / 32][HMMB_MAX_STATES];
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *Bw = Bt + (size_t)w * M * N;
    for (int j = 0; j < N; ++j) {
        double s = 0.0;
        for (int k = tid; k < M; k += RED_THREADS) s += exit_value(Bw[(size_t)k * N + j], finalize);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) sPart[warp][j] = s;
    }
    __syncthreads();
    if (tid < N) {
        double s = 0.0;
        for (int q = 0; q < RED_THREADS / 32; ++q) s += sPart[q][tid];
        sDen[tid] = s;
    }
    __syncthreads();
    for (int e = tid; e < M * N; e += RED_THREADS) {
        const int k = e / N, j = e - k * N;
        double b = exit_value(Bw[e], finalize);
        if (finalize && sDen[j] > 0.0) b = b / sDen[j];
        B_out[((size_t)w * N + j) * M + k] = b;
    }
    if (tid < N) {
        const int i = tid;
        double rs = 0.0;
        for (int j = 0; j < N; ++j) rs += exit_value(A[((size_t)w * N + i) * N + j], finalize);
        for (int j = 0; j < N; ++j) {
            double v = exit_value(A[((size_t)w * N + i) * N + j], finalize);
            if (finalize && rs > 0.0) v = v / rs;
            A_out[((size_t)w * N + i) * N + j] = v;
        }
        double ps = 0.0;
        for (int j = 0; j < N; ++j) ps += exit_value(pi[(size_t)w * N + j], finalize);
        double p = exit_value(pi[(size_t)w * N + i], finalize);
        if (finalize) p = p / ps;  // no zero guard in the reference either (:529)
        pi_out[(size_t)w * N + i] = p;
    }
}

}  // namespace hmmb
