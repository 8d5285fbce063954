// fp64 pipe probe for B200: latency of a dependent DADD/DMUL chain and throughput with ILP x warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_probe fp64_probe.cu && ./fp64_probe
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, double ca, double fb, int iters) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-9 + i;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = __dadd_rn(__dmul_rn(ca, v[i]), fb);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)(t1 - t0) * 0;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (double)(t1 - t0);
}
template <int ILP>
void run(int warps_per_sm, int iters) {
    double* d;
    cudaMalloc(&d, sizeof(double) * 148 * 1024);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    chain<ILP><<<148, warps_per_sm * 32>>>(d, 0.999, 1e-4, 10);
    cudaEventRecord(a);
    chain<ILP><<<148, warps_per_sm * 32>>>(d, 0.999, 1e-4, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    double clk;
    cudaMemcpy(&clk, d, 8, cudaMemcpyDeviceToHost);
    double ops = 2.0 * ILP * iters * warps_per_sm * 32.0 * 148;
    printf("ILP %d warps/SM %2d: %.3f ms, %.1f cycles/iter (2 dependent ops), %.2f Gop/s, %.2f lanes/clk/SM\n", ILP, warps_per_sm, ms,
           clk / iters, ops / ms * 1e-6, ops / (clk * 148));
    cudaFree(d);
}
int main() {
    for (int w : {1, 4, 8, 16, 32}) run<1>(w, 20000);
    for (int w : {1, 4, 8, 16, 32}) run<2>(w, 20000);
    for (int w : {4, 16, 32}) run<4>(w, 20000);
    return 0;
}
