// Does programmatic dependent launch (griddepcontrol) shorten a captured chain of small dependent kernels on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_probe pdl_probe.cu && ./pdl_probe
#include <cstdio>
#include <cuda_runtime.h>
template <bool PDL>
__global__ void __launch_bounds__(256) step(float* buf, int n, int iters) {
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float v = buf[i];
        for (int k = 0; k < iters; ++k) v = fmaf(v, 1.0001f, 0.5f);
        buf[(i + 1) % n] = v;      // depends on the previous kernel's writes
    }
}
template <bool PDL>
float run(float* d, int n, int blocks, int chain, int iters) {
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    for (int k = 0; k < chain; ++k) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = PDL ? 1 : 0;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, step<PDL>, d, n, iters);
    }
    cudaError_t e = cudaStreamEndCapture(st, &g);
    if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1; }
    e = cudaGraphInstantiate(&ge, g, 0);
    if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 5; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(a, st);
    for (int w = 0; w < 20; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms * 1e3f / (20 * chain);
}
int main() {
    float* d;
    const int n = 148 * 256 * 4;
    cudaMalloc(&d, n * 4);
    cudaMemset(d, 0, n * 4);
    for (int blocks : {8, 48, 592}) {
        for (int iters : {100, 4000}) {
            float t0 = run<false>(d, n, blocks, 32, iters), t1 = run<true>(d, n, blocks, 32, iters);
            printf("blocks %3d iters %4d: %.2f us/kernel plain, %.2f us/kernel with PDL\n", blocks, iters, t0, t1);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
