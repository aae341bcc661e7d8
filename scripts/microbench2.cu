// Micro-benchmarks that size the left-to-right N = 16 Baum-Welch kernels (thread = sequence):
// how fast can 32 lanes, each owning a RANDOM 128-byte row of a [1024][16] fp64 table, (a) add
// a 16-vector to their row and (b) read their row, per SM?  Results in DESIGN.md.
//   upd_lock   shared table, per-row spin lock (ATOMS.CAS.32) + swizzled LDS.128/STS.128
//   upd_cas    shared table, 16 x atomicAdd(double) (CAS loop)
//   upd_red    global table (one per CTA, L2 resident), 16 x RED.f64 per lane
//   upd_tma    global table, one cp.reduce.async.bulk .add.f64 of 128 B per lane (row staged in smem)
//   get_ldg    row read with 8 x LDG.128 from the global table
//   get_lds    row read with 8 x LDS.128 from a swizzled shared table
//   get_bulk   row fetched by one cp.async.bulk (128 B) per lane into padded staging, then LDS.128
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/microbench2 scripts/microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__);return 1;}}while(0)

constexpr int M = 1024, NS = 16, CPR = NS / 2, THREADS = 256, WARPS = THREADS / 32;
constexpr int ROWPAD = 144;  // staging row stride: lane l chunk c -> slot (l + c) mod 8, conflict-free per quarter-warp

__device__ __forceinline__ unsigned lcg(unsigned &x) { x = x * 1664525u + 1013904223u; return (x >> 10) & (M - 1); }
__device__ __forceinline__ int swz(unsigned row, int c) { return (int)row * CPR + (c ^ (int)(row & 7u)); }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(THREADS, 1) k_upd_lock(double *out, int iters) {
    extern __shared__ double2 tab[];
    unsigned *locks = reinterpret_cast<unsigned *>(tab + M * CPR);
    for (int i = threadIdx.x; i < M * CPR; i += THREADS) tab[i] = make_double2(0, 0);
    for (int i = threadIdx.x; i < M; i += THREADS) locks[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double g = 1.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
        bool done = false;
        while (!done) {
            if (atomicCAS(&locks[sym], 0u, 1u) == 0u) {
                __threadfence_block();
#pragma unroll
                for (int c = 0; c < CPR; ++c) {
                    double2 v = tab[swz(sym, c)];
                    v.x += g; v.y += g;
                    tab[swz(sym, c)] = v;
                }
                __threadfence_block();
                *reinterpret_cast<volatile unsigned *>(&locks[sym]) = 0u;
                done = true;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = tab[0].x;
}

__global__ void __launch_bounds__(THREADS, 1) k_upd_cas(double *out, int iters) {
    extern __shared__ double2 tab[];
    double *t = reinterpret_cast<double *>(tab);
    for (int i = threadIdx.x; i < M * NS; i += THREADS) t[i] = 0;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double g = 1.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
#pragma unroll
        for (int j = 0; j < NS; ++j) atomicAdd(&t[sym * NS + ((j + 2 * (sym & 7)) & 15)], g);
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t[0];
}

__global__ void __launch_bounds__(THREADS, 1) k_upd_red(double *gtab, int iters) {
    double *t = gtab + (size_t)blockIdx.x * M * NS;
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double g = 1.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
#pragma unroll
        for (int j = 0; j < NS; ++j) atomicAdd(&t[sym * NS + j], g);
    }
}

// lanes = states: 16 lanes add one row (coalesced 128-byte RED), two rows per warp instruction
__global__ void __launch_bounds__(THREADS, 1) k_upd_red_row(double *gtab, int iters) {
    double *t = gtab + (size_t)blockIdx.x * M * NS;
    unsigned x = (threadIdx.x >> 4) * 2654435761u + blockIdx.x * 40503u;
    const int j = threadIdx.x & 15;
    double g = 1.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
        atomicAdd(&t[sym * NS + j], g);
    }
}

__global__ void __launch_bounds__(THREADS, 1) k_upd_tma(double *gtab, int iters) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // two staging buffers per warp, [32 lanes][ROWPAD bytes]
    unsigned char *st0 = stage_raw + (size_t)warp * 2 * 32 * ROWPAD;
    double *t = gtab + (size_t)blockIdx.x * M * NS;
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double g = 1.0 + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
        unsigned char *mine = st0 + (size_t)(it & 1) * 32 * ROWPAD + (size_t)lane * ROWPAD;
        // the buffer used two iterations ago must have been read by the async proxy
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
#pragma unroll
        for (int c = 0; c < CPR; ++c) reinterpret_cast<double2 *>(mine)[c] = make_double2(g, g);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], 128;"
                     :: "l"(t + (size_t)sym * NS), "r"(smem_u32(mine)) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1) k_get_ldg(const double *gtab, double *out, int iters) {
    const double2 *t = reinterpret_cast<const double2 *>(gtab + (size_t)blockIdx.x * M * NS);
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
#pragma unroll
        for (int c = 0; c < CPR; ++c) { const double2 v = __ldg(t + sym * CPR + c); acc += v.x + v.y; }
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(THREADS, 1) k_get_lds(const double *gtab, double *out, int iters) {
    extern __shared__ double2 tab[];
    const double2 *t = reinterpret_cast<const double2 *>(gtab + (size_t)blockIdx.x * M * NS);
    for (int i = threadIdx.x; i < M * CPR; i += THREADS) tab[swz(i / CPR, i % CPR)] = t[i];
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        const unsigned sym = lcg(x);
#pragma unroll
        for (int c = 0; c < CPR; ++c) { const double2 v = tab[swz(sym, c)]; acc += v.x + v.y; }
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(THREADS, 1) k_get_bulk(const double *gtab, double *out, int iters) {
    extern __shared__ __align__(128) unsigned char stage_raw[];
    __shared__ __align__(8) unsigned long long bars[WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *st0 = stage_raw + (size_t)warp * 2 * 32 * ROWPAD;
    const double *t = gtab + (size_t)blockIdx.x * M * NS;
    if (lane == 0) {
        for (int b = 0; b < 2; ++b)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" :: "r"(smem_u32(&bars[warp][b])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    double acc = 0.0;
    auto issue = [&](int buf, unsigned sym) {
        const unsigned bar = smem_u32(&bars[warp][buf]);
        const unsigned dst = smem_u32(st0 + (size_t)buf * 32 * ROWPAD + (size_t)lane * ROWPAD);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 128;" :: "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 128, [%2];"
                     :: "r"(dst), "l"(t + (size_t)sym * NS), "r"(bar) : "memory");
    };
    issue(0, lcg(x));
    for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (it + 1 < iters) issue(buf ^ 1, lcg(x));
        const unsigned bar = smem_u32(&bars[warp][buf]);
        const unsigned parity = (it >> 1) & 1;
        asm volatile(
            "{\n.reg .pred p;\nWAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@!p bra WAIT_%=;\n}\n" :: "r"(bar), "r"(parity) : "memory");
        const double2 *mine = reinterpret_cast<const double2 *>(st0 + (size_t)buf * 32 * ROWPAD + (size_t)lane * ROWPAD);
#pragma unroll
        for (int c = 0; c < CPR; ++c) { const double2 v = mine[c]; acc += v.x + v.y; }
        __syncwarp();
    }
    out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0)); const int sm = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz; %d lanes/CTA, 1 CTA/SM, row = %d fp64 (128 B), table %d rows\n", p.name, sm, p.clockRate, THREADS, NS, M);
    double *gt, *out; CK(cudaMalloc(&gt, (size_t)sm * M * NS * 8)); CK(cudaMemset(gt, 0, (size_t)sm * M * NS * 8));
    CK(cudaMalloc(&out, (size_t)sm * THREADS * 8));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); float ms;
    const int iters = 2000;
    const int smem_tab = M * NS * 8 + M * 4, smem_stage = WARPS * 2 * 32 * ROWPAD;
    CK(cudaFuncSetAttribute(k_upd_lock, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tab));
    CK(cudaFuncSetAttribute(k_upd_cas, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tab));
    CK(cudaFuncSetAttribute(k_get_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tab));
    CK(cudaFuncSetAttribute(k_upd_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_stage));
    CK(cudaFuncSetAttribute(k_get_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_stage));
#define RUN(name, rows_per_cta, ...)                                                                        \
    for (int rep = 0; rep < 2; ++rep) {                                                                     \
        cudaEventRecord(a); __VA_ARGS__; cudaEventRecord(b); CK(cudaEventSynchronize(b)); CK(cudaGetLastError()); \
        cudaEventElapsedTime(&ms, a, b);                                                                    \
        if (rep) printf("%-12s %8.3f ms  %7.2f G rows/s chip  %6.2f clk per row per SM\n", name, ms,       \
                        (double)sm * (rows_per_cta) / ms / 1e6, ms * 1e-3 * p.clockRate * 1e3 / (double)(rows_per_cta)); \
    }
    const double rows = (double)THREADS * iters;
    RUN("upd_lock", rows, (k_upd_lock<<<sm, THREADS, smem_tab>>>(out, iters)));
    RUN("upd_cas", rows, (k_upd_cas<<<sm, THREADS, smem_tab>>>(out, iters)));
    RUN("upd_red", rows, (k_upd_red<<<sm, THREADS>>>(gt, iters)));
    RUN("upd_red_row", rows, (k_upd_red_row<<<sm, THREADS>>>(gt, iters * 16)));
    RUN("upd_tma", rows, (k_upd_tma<<<sm, THREADS, smem_stage>>>(gt, iters)));
    RUN("get_ldg", rows, (k_get_ldg<<<sm, THREADS>>>(gt, out, iters)));
    RUN("get_lds", rows, (k_get_lds<<<sm, THREADS, smem_tab>>>(gt, out, iters)));
    RUN("get_bulk", rows, (k_get_bulk<<<sm, THREADS, smem_stage>>>(gt, out, iters)));
    return 0;
}
