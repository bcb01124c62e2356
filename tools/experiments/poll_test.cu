// Does a kernel polling a flag in device memory see a store made by a kernel of another stream of the same process?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void poller(unsigned* flag, unsigned* out, int mode) {
  unsigned spins = 0, v = 0;
  while (true) {
    if (mode == 0) v = *(volatile unsigned*)flag;
    else if (mode == 1) asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    else if (mode == 2) v = atomicAdd_system(flag, 0u);
    else asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v) break;
    __nanosleep(200);
    if (++spins > (1u << 22)) break;
  }
  out[0] = v; out[1] = spins;
}
__global__ void writer(unsigned* flag, int mode) {
  if (mode == 0) *(volatile unsigned*)flag = 1u;
  else if (mode == 1) asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(1u) : "memory");
  else atomicMax_system(flag, 1u);
  __threadfence_system();
}
int main() {
  cudaStream_t s0, s1;
  cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  unsigned *flag, *out, h[2];
  cudaMalloc(&flag, 256); cudaMalloc(&out, 8);
  unsigned* pinned; cudaHostAlloc((void**)&pinned, 64, cudaHostAllocDefault);
  cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventBlockingSync);
  for (int follow = 0; follow < 4; follow++)
  for (int pm = 1; pm < 3; pm++)
    for (int wm = 1; wm < 3; wm++) {
      cudaMemset(flag, 0, 256); cudaMemset(out, 0, 8);
      cudaDeviceSynchronize();
      poller<<<1, 1, 0, s0>>>(flag, out, pm);
      if (follow >= 1) writer<<<1, 1, 0, s0>>>(flag + 32, wm);                       // a kernel queued behind the poller
      if (follow >= 2) cudaMemcpyAsync(pinned, out, 8, cudaMemcpyDeviceToHost, s0);  // ... and a copy behind that
      if (follow >= 3) cudaEventRecord(ev, s0);
      writer<<<1, 1, 0, s1>>>(flag, wm);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, 8, cudaMemcpyDeviceToHost);
      printf("follow %d poll mode %d write mode %d: saw %u after %u spins (%s)\n", follow, pm, wm, h[0], h[1], cudaGetErrorString(e));
    }
  return 0;
}
