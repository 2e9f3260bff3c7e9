// Microbenchmark: per-kernel cost of a chain of N dependent tiny kernels inside a CUDA graph, with and without
// programmatic dependent launch (griddepcontrol.wait at kernel entry).  nvcc -arch=sm_100a -o pdl_gap pdl_gap.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void step_kernel(float* x, int pdl) {
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  x[i] = x[i] * 1.0001f + 1.f;
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
static float run(int pdl, int n_kernels, int ctas) {
  float* x;
  cudaMalloc(&x, ctas * 256 * sizeof(float));
  cudaMemset(x, 0, ctas * 256 * sizeof(float));
  cudaStream_t s;
  cudaStreamCreate(&s);
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int k = 0; k < n_kernels; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, step_kernel, x, pdl);
  }
  cudaError_t e = cudaStreamEndCapture(s, &g);
  if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1.f; }
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, s);
  cudaStreamSynchronize(s);
  cudaEventRecord(a, s);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, s);
  cudaEventRecord(b, s);
  cudaStreamSynchronize(s);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  float h;
  cudaMemcpy(&h, x, 4, cudaMemcpyDeviceToHost);
  printf("pdl=%d ctas=%d: %.3f us per kernel (x[0]=%.1f, err=%s)\n", pdl, ctas, ms * 1000.f / (10.f * n_kernels), h,
         cudaGetErrorString(cudaGetLastError()));
  return ms;
}
int main() {
  for (int ctas : {1, 148, 592}) { run(0, 400, ctas); run(1, 400, ctas); }
  return 0;
}
