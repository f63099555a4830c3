// What does a kernel launch cost inside a CUDA graph on this box, as a function of the launch configuration?
// Graphs of 40 dependent launches; reports microseconds per launch.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

__global__ void k_empty(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
__global__ void k_tmap(const __grid_constant__ CUtensorMap a, const __grid_constant__ CUtensorMap b, int* p) {
  if (p && threadIdx.x == 9999) *p = reinterpret_cast<const int*>(&a)[0] + reinterpret_cast<const int*>(&b)[0];
}
__global__ void k_tmem(int* p) {
  __shared__ uint32_t slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
  if (p && threadIdx.x == 9999) *p = 1;
}
__global__ void k_pdl(int* p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p && threadIdx.x == 9999) *p = 1;
}

template <typename F>
static float graph_time(cudaStream_t st, F launch, int n = 40, int reps = 20) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < n; ++i) launch();
  cudaStreamEndCapture(st, &g);
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1; }
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, st);
  for (int r = 0; r < reps; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1e3f / (n * reps);
}

int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  int* d; cudaMalloc(&d, 4);
  CUtensorMap ta, tb; memset(&ta, 0, sizeof(ta)); memset(&tb, 0, sizeof(tb));
  cudaFuncSetAttribute(k_empty, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tmap, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("empty 18x256, smem 0          : %6.2f us/launch\n", graph_time(st, [&] { k_empty<<<18, 256, 0, st>>>(d); }));
  printf("empty 18x384, smem 0          : %6.2f us/launch\n", graph_time(st, [&] { k_empty<<<18, 384, 0, st>>>(d); }));
  printf("empty 148x384, smem 0         : %6.2f us/launch\n", graph_time(st, [&] { k_empty<<<148, 384, 0, st>>>(d); }));
  printf("empty 18x384, smem 100KB      : %6.2f us/launch\n", graph_time(st, [&] { k_empty<<<18, 384, 100 * 1024, st>>>(d); }));
  printf("empty 18x384, smem 200KB      : %6.2f us/launch\n", graph_time(st, [&] { k_empty<<<18, 384, 200 * 1024, st>>>(d); }));
  printf("empty 148x384, smem 200KB     : %6.2f us/launch\n", graph_time(st, [&] { k_empty<<<148, 384, 200 * 1024, st>>>(d); }));
  printf("alternate smem 200KB / 0      : %6.2f us/launch\n", graph_time(st, [&] { static int i = 0; if ((i++) & 1) k_empty<<<18, 384, 200 * 1024, st>>>(d); else k_empty<<<8, 256, 0, st>>>(d); }));
  printf("2 tensormap params, smem 0    : %6.2f us/launch\n", graph_time(st, [&] { k_tmap<<<18, 384, 0, st>>>(ta, tb, d); }));
  printf("2 tensormap params, smem 200KB: %6.2f us/launch\n", graph_time(st, [&] { k_tmap<<<18, 384, 200 * 1024, st>>>(ta, tb, d); }));
  printf("tmem alloc/dealloc, smem 0    : %6.2f us/launch\n", graph_time(st, [&] { k_tmem<<<18, 384, 0, st>>>(d); }));
  printf("tmem alloc/dealloc, smem 200KB: %6.2f us/launch\n", graph_time(st, [&] { k_tmem<<<18, 384, 200 * 1024, st>>>(d); }));
  auto pdl_launch = [&](int smem) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(18); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, k_pdl, d);
  };
  cudaFuncSetAttribute(k_pdl, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("PDL chain, smem 0             : %6.2f us/launch\n", graph_time(st, [&] { pdl_launch(0); }));
  printf("PDL chain, smem 200KB         : %6.2f us/launch\n", graph_time(st, [&] { pdl_launch(200 * 1024); }));
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
