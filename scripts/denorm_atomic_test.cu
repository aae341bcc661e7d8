// Do fp64 atomics keep denormals?  (global RED / ATOM, shared CAS loop, plain add)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *g, double tiny) {
    __shared__ double s[2];
    if (threadIdx.x == 0) { s[0] = 0.0; s[1] = 1e-310; }
    __syncthreads();
    atomicAdd(g + 0, tiny);                 // RED, 0 + denormal, 32 times
    atomicAdd(g + 1, tiny);                 // 1e-310 + denormal
    double old = atomicAdd(g + 2, tiny);    // ATOM (value used)
    if (old < 0) g[7] = old;
    atomicAdd(s + 0, tiny);
    atomicAdd(s + 1, tiny);
    __syncthreads();
    if (threadIdx.x == 0) { g[3] = s[0]; g[4] = s[1]; g[5] = 0.0 + tiny; g[6] = 1e-310 + tiny; }
}
int main() {
    double h[8] = {0.0, 1e-310, 0.0, -1, -1, -1, -1, 0}, *d;
    cudaMalloc(&d, sizeof h); cudaMemcpy(d, h, sizeof h, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(d, 4.9406564584124654e-324);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("RED 0+32*tiny = %.3e (want 1.58e-322)\nRED 1e-310+32*tiny = %.17e\nATOM 0+32*tiny = %.3e\nshared 0+32*tiny = %.3e\nshared 1e-310+.. = %.17e\nplain 0+tiny = %.3e, 1e-310+tiny = %.17e\n",
           h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    return 0;
}
