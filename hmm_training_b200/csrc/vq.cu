// VQ encode and LBG (binary-split k-means) codebook kernels for sm_100a.
//
// Replaces (reference paths relative to the reference root):
//   get_observations ........ HMM/hmm_training.py:82-120
//   createCodeVector ........ CodeVector/codevector_functions.py:442-531
//   new_epsilon_centroids ... CodeVector/codevector_functions.py:383-411
//   new_adjust_centroids .... CodeVector/codevector_functions.py:414-439
//
// Arithmetic contract (bit-exact indices): the reference computes
// np.linalg.norm(x[1:] - c[1:]) = sqrt(dot(d, d)); with the image's numpy/OpenBLAS the
// 12-term dot is the sequential FMA chain acc = fma(d_i, d_i, acc), i = 1..12.  The kernel
// evaluates exactly that chain in fp64 (__dsub_rn + __fma_rn, no re-association), compares
// with strict '<' on the square root (only evaluated when the squared distance improves),
// so the lowest index wins ties exactly as at hmm_training.py:112.
//
// Roofline: 12 DADD + 12 DFMA per (frame, centroid) pair -> FP64-pipe bound
// (AI ~ 85 flop/B); HBM traffic is 104 B in + 4 B out per frame.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace hmmb {

constexpr int VQ_THREADS = 256;
constexpr int VQ_TILE = 512;  // centroids per shared-memory tile: 512 * 12 * 8 B = 48 KB
constexpr int VQ_D = 12;      // dims 1..12 take part in the distance
#ifndef VQ_ILP
#define VQ_ILP 4              // centroids in flight per thread (4 / 8 and one-batch-per-CTA grids all measured 0.47 ms:
                              // the kernel sits at ~78 % of the fp64 issue rate counting the 25 DP instructions per pair)
#endif
constexpr int ACC_W = 14;     // per-centroid accumulator row: 13 sums + count
constexpr int VQ_PRIV_K = 32; // up to this many centroids every warp keeps a private accumulator table

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MODE 0: encode only.  MODE 1: LBG pass = encode + per-centroid 13-dim sums, counts and
// the sum of winning distances, reduced in shared memory and flushed once per CTA.
template <int MODE>
__global__ void __launch_bounds__(VQ_THREADS)
k_vq_assign(const double *__restrict__ X, int64_t F, const double *__restrict__ C, int K,
            int32_t *__restrict__ idx_out, double *__restrict__ dist_out, double *__restrict__ accum,
            int smem_accum, const int *__restrict__ skip) {
    // a Lloyd pass queued speculatively after the generation has converged (k_lbg_check) is a no-op
    if (skip && *skip) return;
    extern __shared__ double smem[];
    double *sC = smem;                                          // [tile][12]
    double *sAcc = smem + (size_t)min(K, VQ_TILE) * VQ_D;       // [K][14] (MODE 1, if smem_accum)
    __shared__ double sRed[VQ_THREADS / 32];

    const int tid = threadIdx.x, lane = tid & 31;
    // Few centroids (Lloyd passes of the first LBG generations): all lanes hit the same handful of rows, so
    // shared atomics serialise.  Every warp then owns a private [K][14] table: the lanes park their frame and
    // its key in shared memory and lanes 0..13 (one per accumulator column) walk the 32 frames, adding column
    // d of frame l to row key[l] — no atomics, no shuffles, cost independent of K.
    const bool priv = MODE == 1 && smem_accum == 2;
    double *wAcc = sAcc + (size_t)(tid >> 5) * K * ACC_W;                                       // [K][14] per warp
    double *wX = sAcc + (size_t)(VQ_THREADS / 32) * K * ACC_W + (size_t)(tid >> 5) * 32 * 13;   // [32][13] per warp
    int *wKey = reinterpret_cast<int *>(sAcc + (size_t)(VQ_THREADS / 32) * (K * ACC_W + 32 * 13)) + (tid >> 5) * 32;
    if (MODE == 1 && smem_accum) {
        const int n = priv ? (VQ_THREADS / 32) * K * ACC_W : K * ACC_W;
        for (int e = tid; e < n; e += VQ_THREADS) sAcc[e] = 0.0;
    }
    double dist_local = 0.0;
    const int ntiles = (K + VQ_TILE - 1) / VQ_TILE;
    const int64_t nbatch = (F + VQ_THREADS - 1) / VQ_THREADS;

    for (int64_t batch = blockIdx.x; batch < nbatch; batch += gridDim.x) {
        const int64_t f = batch * VQ_THREADS + tid;
        const bool valid = f < F;
        double x[13];
#pragma unroll
        for (int d = 0; d < 13; ++d) x[d] = valid ? __ldg(X + f * 13 + d) : 0.0;

        double best_d2 = pos_inf(), best_s = pos_inf();
        int best = 0;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int k0 = tile * VQ_TILE;
            const int kt = min(VQ_TILE, K - k0);
            if (ntiles > 1 || batch == blockIdx.x) {
                __syncthreads();
                for (int e = tid; e < kt * VQ_D; e += VQ_THREADS) {
                    int k = e / VQ_D, d = e - k * VQ_D;
                    sC[e] = __ldg(C + (size_t)(k0 + k) * 13 + 1 + d);
                }
                __syncthreads();
            }
            int k = 0;
            // VQ_ILP centroids in flight per thread: independent sub -> fma chains keep the fp64 pipe fed
            for (; k + VQ_ILP <= kt; k += VQ_ILP) {
                const double2 *c0 = reinterpret_cast<const double2 *>(sC + (size_t)k * VQ_D);
                double a[VQ_ILP];
#pragma unroll
                for (int u = 0; u < VQ_ILP; ++u) a[u] = 0.0;
#pragma unroll
                for (int q = 0; q < VQ_D / 2; ++q) {
#pragma unroll
                    for (int u = 0; u < VQ_ILP; ++u) {
                        const double2 p = c0[q + u * (VQ_D / 2)];
                        double v;
                        v = __dsub_rn(x[1 + 2 * q], p.x); a[u] = __fma_rn(v, v, a[u]);
                        v = __dsub_rn(x[2 + 2 * q], p.y); a[u] = __fma_rn(v, v, a[u]);
                    }
                }
                // in index order; sqrt only when the squared distance improves
#pragma unroll
                for (int u = 0; u < VQ_ILP; ++u)
                    if (a[u] < best_d2) { double s = sqrt(a[u]); if (s < best_s) { best_s = s; best_d2 = a[u]; best = k0 + k + u; } }
            }
            for (; k < kt; ++k) {
                const double *c = sC + (size_t)k * VQ_D;
                double a = 0.0;
#pragma unroll
                for (int d = 0; d < VQ_D; ++d) {
                    double v = __dsub_rn(x[1 + d], c[d]);
                    a = __fma_rn(v, v, a);
                }
                if (a < best_d2) { double s = sqrt(a); if (s < best_s) { best_s = s; best_d2 = a; best = k0 + k; } }
            }
        }
        if (valid) {
            if (idx_out) idx_out[f] = best;
            if (dist_out) dist_out[f] = best_s;
        }
        if (MODE == 1) {
            if (valid) dist_local += best_s;
            double *acc = smem_accum ? sAcc : accum;
            if (priv) {
#pragma unroll
                for (int d = 0; d < 13; ++d) wX[lane * 13 + d] = x[d];
                wKey[lane] = valid ? best : -1;
                __syncwarp();
                if (lane < ACC_W) {
                    for (int l = 0; l < 32; ++l) {
                        const int k = wKey[l];
                        if (k >= 0) wAcc[k * ACC_W + lane] += (lane < 13) ? wX[l * 13 + lane] : 1.0;
                    }
                }
                __syncwarp();
            } else if (K <= 16) {
                // few centroids: every lane hits the same handful of rows, so reduce per
                // key across the warp with shuffles and issue one atomic per (key, dim).
                unsigned remaining = __ballot_sync(0xffffffffu, valid);
                while (remaining) {
                    int leader = __ffs(remaining) - 1;
                    int key = __shfl_sync(0xffffffffu, best, leader);
                    bool mine = valid && best == key;
                    unsigned m = __ballot_sync(0xffffffffu, mine);
                    remaining &= ~m;
#pragma unroll
                    for (int d = 0; d < 13; ++d) {
                        double v = warp_sum(mine ? x[d] : 0.0);
                        if (lane == 0) atomicAdd(acc + key * ACC_W + d, v);
                    }
                    if (lane == 0) atomicAdd(acc + key * ACC_W + 13, (double)__popc(m));
                }
            } else if (valid) {
#pragma unroll
                for (int d = 0; d < 13; ++d) atomicAdd(acc + best * ACC_W + d, x[d]);
                atomicAdd(acc + best * ACC_W + 13, 1.0);
            }
        }
    }
    if (MODE == 1) {
        double v = warp_sum(dist_local);
        if (lane == 0) sRed[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
            for (int w = 0; w < VQ_THREADS / 32; ++w) s += sRed[w];
            atomicAdd(accum + (size_t)K * ACC_W, s);
        }
        if (priv) {
            __syncthreads();
            for (int e = tid; e < K * ACC_W; e += VQ_THREADS) {
                double a = 0.0;
#pragma unroll
                for (int w = 0; w < VQ_THREADS / 32; ++w) a += sAcc[(size_t)w * K * ACC_W + e];
                if (a != 0.0) atomicAdd(accum + e, a);
            }
        } else if (smem_accum) {
            for (int e = tid; e < K * ACC_W; e += VQ_THREADS) {
                double a = sAcc[e];
                if (a != 0.0) atomicAdd(accum + e, a);
            }
        }
    }
}

// new_adjust_centroids: mean of the assigned frames (all 13 dims), zeros(13) if empty.
// Convergence state of one LBG generation, kept on the device so that Lloyd passes can be queued
// several at a time without a host round trip per pass (codevector_functions.py:475-476, :485, :509-510).
struct LbgState {
    double prev, gd;
    int it, done;
};
__global__ void k_lbg_reset(LbgState *st) {
    st->prev = 0.0;
    st->gd = 0.0;
    st->it = 0;
    st->done = 0;
}
// after the centroid update of a pass: count it, compare the summed distance with the previous pass
__global__ void k_lbg_check(const double *__restrict__ gdp, LbgState *st, double eps) {
    if (st->done) return;
    const double gd = *gdp;
    const double diff = fabs(st->prev - gd);
    st->it += 1;
    st->prev = gd;
    st->gd = gd;
    if (!(diff > eps)) st->done = 1;  // loop condition `while diff > eps` (:485)
}

__global__ void k_lbg_update(const double *__restrict__ accum, int K, double *__restrict__ C,
                             const LbgState *__restrict__ st) {
    if (st && st->done) return;  // (the flag only changes in k_lbg_check, a separate launch)
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * 13) return;
    int k = e / 13, d = e - k * 13;
    double n = accum[k * ACC_W + 13];
    C[e] = n > 0.0 ? accum[k * ACC_W + d] / n : 0.0;
}

// new_epsilon_centroids: centroid i -> 2i = c * 1.001, 2i+1 = c * 0.999 (all 13 dims).
__global__ void k_lbg_split(const double *__restrict__ C, int K, double *__restrict__ C2) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= K * 13) return;
    int k = e / 13, d = e - k * 13;
    double v = C[e];
    C2[(size_t)(2 * k) * 13 + d] = v * 1.001;
    C2[(size_t)(2 * k + 1) * 13 + d] = v * 0.999;
}

// smem_accum: 0 = accumulate straight into global memory, 1 = one shared [K][14] table per CTA,
// 2 = one private table per warp (K <= VQ_PRIV_K)
static size_t vq_smem_bytes(int K, int mode, int *smem_accum) {
    size_t tile = (size_t)std::min(K, VQ_TILE) * VQ_D * sizeof(double);
    size_t acc = (size_t)K * ACC_W * sizeof(double);
    *smem_accum = 0;
    if (mode == 1 && K <= VQ_PRIV_K && !getenv("HMMB_LBG_NO_PRIVATE")) {
        *smem_accum = 2;
        const size_t warps = VQ_THREADS / 32;
        return tile + warps * (acc + 32 * 13 * sizeof(double)) + warps * 32 * sizeof(int);
    }
    if (mode == 1 && tile + acc <= 200 * 1024) {
        *smem_accum = 1;
        return tile + acc;
    }
    return tile;
}

static int vq_launch(int mode, const double *dX, int64_t F, const double *dC, int K, int32_t *d_idx,
                     double *d_dist, double *d_accum, const int *d_skip = nullptr) {
    Ctx &c = ctx();
    int smem_accum = 0;
    size_t smem = vq_smem_bytes(K, mode, &smem_accum);
    int64_t nbatch = (F + VQ_THREADS - 1) / VQ_THREADS;
    if (nbatch == 0) return HMMB_OK;
    // persistent-style grid: a multiple of the SM count, each CTA strides over frame batches
    int per_sm = smem > 100 * 1024 ? 1 : (smem > 60 * 1024 ? 3 : 4);
    int64_t grid = (int64_t)c.sm_count * per_sm * (mode == 1 ? 1 : 4);
    if (grid > nbatch) grid = nbatch;
    if (mode == 0) {
        HMMB_CUDA(cudaFuncSetAttribute(k_vq_assign<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HMMB_LAUNCH("vq_encode", k_vq_assign<0>, (unsigned)grid, VQ_THREADS, smem, dX, F, dC, K, d_idx, d_dist,
                    d_accum, smem_accum, d_skip);
    } else {
        HMMB_CUDA(cudaFuncSetAttribute(k_vq_assign<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        HMMB_LAUNCH("lbg_assign", k_vq_assign<1>, (unsigned)grid, VQ_THREADS, smem, dX, F, dC, K, d_idx, d_dist,
                    d_accum, smem_accum, d_skip);
    }
    return HMMB_OK;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { dev_free(p); }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace hmmb

using namespace hmmb;

extern "C" {

int hmmb_vq_encode_dev(const double *dX, int64_t F, const double *dC, int K, int32_t *d_idx, double *d_dist) {
    HMMB_TRY(require_init());
    if (F < 0 || K <= 0 || !dC || (F > 0 && (!dX || !d_idx))) {
        set_error("hmmb_vq_encode_dev: bad arguments (F=%lld, K=%d)", (long long)F, K);
        return HMMB_ERR_ARG;
    }
    return vq_launch(0, dX, F, dC, K, d_idx, d_dist, nullptr);
}

int hmmb_vq_encode(const double *X, int64_t F, const double *C, int K, int32_t *idx_out) {
    HMMB_TRY(require_init());
    if (F < 0 || K <= 0 || !C || (F > 0 && (!X || !idx_out))) {
        set_error("hmmb_vq_encode: bad arguments (F=%lld, K=%d)", (long long)F, K);
        return HMMB_ERR_ARG;
    }
    if (F == 0) return HMMB_OK;
    Ctx &c = ctx();
    DevBuf dX, dC, dI;
    HMMB_TRY(dev_alloc(&dX.p, (size_t)F * 13 * sizeof(double)));
    HMMB_TRY(dev_alloc(&dC.p, (size_t)K * 13 * sizeof(double)));
    HMMB_TRY(dev_alloc(&dI.p, (size_t)F * sizeof(int32_t)));
    HMMB_CUDA(cudaMemcpyAsync(dC.p, C, (size_t)K * 13 * sizeof(double), cudaMemcpyHostToDevice, c.stream));
    // Frames in pinned host memory go up in chunks on the copy stream and every chunk is encoded as soon
    // as it has landed: the kernel (2 G frames/s) hides behind the PCIe transfer (0.5 G frames/s).
    cudaPointerAttributes attr;
    const bool pinned = cudaPointerGetAttributes(&attr, X) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    const int64_t chunk = 1 << 17;  // 131 072 frames = 13.6 MB
    if (pinned && F >= 2 * chunk) {
        cudaEvent_t fence = event_get();  // recycled device blocks may still be in use on the compute stream
        HMMB_CUDA(cudaEventRecord(fence, c.stream));
        HMMB_CUDA(cudaStreamWaitEvent(c.copy_stream, fence, 0));
        event_put(fence);
        for (int64_t f0 = 0; f0 < F; f0 += chunk) {
            const int64_t n = std::min<int64_t>(chunk, F - f0);
            HMMB_CUDA(cudaMemcpyAsync(dX.as<double>() + f0 * 13, X + f0 * 13, (size_t)n * 13 * sizeof(double),
                                      cudaMemcpyHostToDevice, c.copy_stream));
            cudaEvent_t landed = event_get();
            HMMB_CUDA(cudaEventRecord(landed, c.copy_stream));
            HMMB_CUDA(cudaStreamWaitEvent(c.stream, landed, 0));
            event_put(landed);
            HMMB_TRY(vq_launch(0, dX.as<double>() + f0 * 13, n, dC.as<double>(), K, dI.as<int32_t>() + f0, nullptr, nullptr));
        }
    } else {
        HMMB_TRY(h2d_big(dX.p, X, (size_t)F * 13 * sizeof(double), c.stream));  // pageable frames: pinned bounce buffers
        HMMB_TRY(vq_launch(0, dX.as<double>(), F, dC.as<double>(), K, dI.as<int32_t>(), nullptr, nullptr));
    }
    HMMB_TRY(d2h_big(idx_out, dI.p, (size_t)F * sizeof(int32_t), c.stream));
    return HMMB_OK;
}

int hmmb_lbg_fit(const double *X, int64_t F, int x_on_device, int K, int max_iter, double eps, double *C_out,
                 double *gens_out, int32_t *assign_out, int32_t *iters_per_gen, double *gdist_out,
                 hmmb_allreduce_fn allreduce, void *user) {
    HMMB_TRY(require_init());
    if (F <= 0 && !allreduce) {
        set_error("No raw data provided");  // codevector_functions.py:445-446
        return HMMB_ERR_EMPTY;
    }
    if (K <= 0 || !C_out || !gens_out || (F > 0 && !X) || F < 0) {
        set_error("hmmb_lbg_fit: bad arguments (F=%lld, K=%d)", (long long)F, K);
        return HMMB_ERR_ARG;
    }
    Ctx &c = ctx();
    int n_gen = 0;
    while ((2 << n_gen) <= K) ++n_gen;  // floor(log2 K)
    const int Kmax = 1 << (n_gen > 0 ? n_gen : 1);
    DevBuf dXb, dCa, dCb, dI, dAcc;
    const double *dX = X;
    if (!x_on_device && F > 0) {
        HMMB_TRY(dev_alloc(&dXb.p, (size_t)F * 13 * sizeof(double)));
        HMMB_TRY(h2d_big(dXb.p, X, (size_t)F * 13 * sizeof(double), c.stream));
        dX = dXb.as<double>();
    }
    HMMB_TRY(dev_alloc(&dCa.p, (size_t)Kmax * 13 * sizeof(double)));
    HMMB_TRY(dev_alloc(&dCb.p, (size_t)Kmax * 13 * sizeof(double)));
    HMMB_TRY(dev_alloc(&dI.p, (size_t)(F > 0 ? F : 1) * sizeof(int32_t)));
    const size_t acc_n = (size_t)Kmax * ACC_W + 2;
    HMMB_TRY(dev_alloc(&dAcc.p, acc_n * sizeof(double)));
    double *cur = dCa.as<double>(), *nxt = dCb.as<double>();
    double *acc = dAcc.as<double>();
    int32_t *d_idx = dI.as<int32_t>();
    std::vector<double> hC((size_t)Kmax * 13);

    // C0 = mean of all frames (:458-459): one accumulate pass against a single zero centroid
    HMMB_CUDA(cudaMemsetAsync(cur, 0, 13 * sizeof(double), c.stream));
    HMMB_CUDA(cudaMemsetAsync(acc, 0, acc_n * sizeof(double), c.stream));
    HMMB_TRY(vq_launch(1, dX, F, cur, 1, d_idx, nullptr, acc));
    if (allreduce) {
        int rc = allreduce(acc, ACC_W + 1, user);
        if (rc != 0) { set_error("allreduce hook failed (%d)", rc); return HMMB_ERR_CUDA; }
    }
    HMMB_LAUNCH("lbg_update", k_lbg_update, 1, 32, 0, acc, 1, cur, (const LbgState *)nullptr);
    size_t gpos = 0;
    HMMB_CUDA(cudaMemcpyAsync(gens_out, cur, 13 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    gpos += 13;
    HMMB_LAUNCH("lbg_split", k_lbg_split, 1, 32, 0, cur, 1, nxt);  // :469
    { double *t = cur; cur = nxt; nxt = t; }
    int Kg = 2;
    HMMB_CUDA(cudaMemsetAsync(d_idx, 0, (size_t)(F > 0 ? F : 1) * sizeof(int32_t), c.stream));

    DevBuf dState;
    HMMB_TRY(dev_alloc(&dState.p, sizeof(LbgState)));
    LbgState *st = dState.as<LbgState>();
    constexpr int LBG_QUEUE = 8;  // Lloyd passes queued per host round trip
    for (int g = 1; g <= n_gen; ++g) {
        // Lloyd loop of the generation (:475-514).  The convergence test runs on the device (k_lbg_check); passes
        // are queued LBG_QUEUE at a time and the ones behind the converged pass return at once, so the host
        // synchronises once per LBG_QUEUE passes instead of once per pass.
        HMMB_LAUNCH("lbg_update", k_lbg_reset, 1, 1, 0, st);
        double *buf[2] = {cur, nxt};
        LbgState hs{0.0, 0.0, 0, 0};
        int enq = 0;
        while (!hs.done && enq < max_iter) {
            const int n = std::min(LBG_QUEUE, max_iter - enq);
            for (int i = 0; i < n; ++i) {
                const int p = enq + i;
                HMMB_CUDA(cudaMemsetAsync(acc, 0, ((size_t)Kg * ACC_W + 1) * sizeof(double), c.stream));
                HMMB_TRY(vq_launch(1, dX, F, buf[p & 1], Kg, d_idx, nullptr, acc, &st->done));
                if (allreduce) {
                    int rc = allreduce(acc, (int64_t)Kg * ACC_W + 1, user);
                    if (rc != 0) { set_error("allreduce hook failed (%d)", rc); return HMMB_ERR_CUDA; }
                }
                HMMB_LAUNCH("lbg_update", k_lbg_update, (Kg * 13 + 127) / 128, 128, 0, acc, Kg, buf[(p + 1) & 1], st);
                HMMB_LAUNCH("lbg_update", k_lbg_check, 1, 1, 0, acc + (size_t)Kg * ACC_W, st, eps);
            }
            enq += n;
            HMMB_CUDA(cudaMemcpyAsync(&hs, st, sizeof(LbgState), cudaMemcpyDeviceToHost, c.stream));
            HMMB_CUDA(cudaStreamSynchronize(c.stream));
        }
        const int it = hs.it;
        const double gd = hs.gd;
        cur = buf[it & 1];  // written by the last pass that ran
        nxt = buf[(it + 1) & 1];
        if (iters_per_gen) iters_per_gen[g - 1] = it;
        if (gdist_out) gdist_out[g - 1] = gd;
        HMMB_CUDA(cudaMemcpyAsync(gens_out + gpos, cur, (size_t)Kg * 13 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        gpos += (size_t)Kg * 13;
        if (g < n_gen) {  // :520-521
            HMMB_LAUNCH("lbg_split", k_lbg_split, (Kg * 13 + 127) / 128, 128, 0, cur, Kg, nxt);
            { double *t = cur; cur = nxt; nxt = t; }
            Kg *= 2;
        }
    }
    HMMB_CUDA(cudaMemcpyAsync(C_out, cur, (size_t)Kg * 13 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
    if (assign_out && F > 0)
        HMMB_CUDA(cudaMemcpyAsync(assign_out, d_idx, (size_t)F * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
    HMMB_CUDA(cudaStreamSynchronize(c.stream));
    return Kg;
}

}  // extern "C"
