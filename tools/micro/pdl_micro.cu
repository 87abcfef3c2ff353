// Microbenchmark: cost per node of a CUDA graph of N small DEPENDENT kernels, with and without programmatic dependent
// launch (cudaLaunchAttributeProgrammaticStreamSerialization + griddepcontrol).  nvcc -arch=sm_100a -o pdl_micro pdl_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void small_kernel(float* p, int n, int pdl) {
    if (pdl) asm volatile("griddepcontrol.launch_dependents;");
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0001f + 1.0f;
}
static float run(int nodes, int blocks, int pdl, float* d, int n) {
    cudaStream_t st; cudaStreamCreate(&st);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int k = 0; k < nodes; k++) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
        cudaLaunchKernelEx(&cfg, small_kernel, d, n, pdl);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; w++) cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    const int reps = 20;
    for (int r = 0; r < reps; r++) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    return ms * 1000.f / (reps * nodes);
}
int main() {
    const int n = 148 * 256 * 8;
    float* d; cudaMalloc(&d, n * 4); cudaMemset(d, 0, n * 4);
    for (int blocks : {1, 148, 148 * 8}) {
        float a = run(200, blocks, 0, d, n), b = run(200, blocks, 1, d, n);
        printf("graph of 200 dependent kernels, %4d blocks x 256 threads: %.2f us per node plain, %.2f us with PDL\n", blocks, a, b);
    }
    return 0;
}
