// Microbenchmark: raw tcgen05.mma issue rate per kind / shape on one SM and on all SMs (clock64 around a
// chain of N back-to-back MMAs on resident smem operands, one commit at the end).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int KIND> __device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  if (KIND == 0) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc));
  if (KIND == 1) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc));
  if (KIND == 2) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc));
  if (KIND == 3) asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc));
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) bench(uint32_t idesc, int n_mma, int n_acc, int ncols, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_s;
  const uint32_t base = (smem_u32(sm) + 1023) & ~1023u;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)sm)[i] = 0x01010101u * (i & 1);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;");
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tm = tmem_s;
  if (threadIdx.x < 32) {          // whole warp runs the loop (warp-uniform), one elected lane issues
    uint64_t da[4], db[4];
    for (int k = 0; k < 4; ++k) { da[k] = desc_sw128(base + k * 32); db[k] = desc_sw128(base + 32768 + k * 32); }
    const uint32_t tm1 = tm + (n_acc > 1 ? ncols : 0);
    uint32_t elected;
    asm volatile("{.reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P;}" : "=r"(elected));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 8) {
      if (elected) {
#pragma unroll
        for (int k = 0; k < 4; ++k) mma<KIND>(tm, da[k], db[k], idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma<KIND>(tm1, da[k], db[k], idesc, 1u);
      }
    }
    if (elected) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)));
    uint32_t done = 0;
    while (!done) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(done) : "r"(smem_u32(&bar)));
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

static uint32_t idesc(int cfmt, int afmt, int bfmt, int M, int N) {
  return (cfmt << 4) | (afmt << 7) | (bfmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int KIND> void run(const char* name, uint32_t id, int M, int N, int K, int grid) {
  long long* d; cudaMalloc(&d, 8 * 256);
  const int n_mma = 4096, n_acc = 512 / N >= 2 ? 2 : 1;
  cudaFuncSetAttribute(bench<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) bench<KIND><<<grid, 128, 180 * 1024>>>(id, n_mma, n_acc, N, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
  double macs = (double)M * N * K * n_mma;
  printf("%-34s grid=%3d  %-12s clk/MMA=%7.1f  MAC/clk/SM=%8.1f\n", name, grid, cudaGetErrorString(e), (double)mx / n_mma, macs / mx);
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<0>("f16  M128 N128 K16 (fp32 acc)", idesc(1, 0, 0, 128, 128), 128, 128, 16, grid);
    run<0>("f16  M128 N256 K16 (fp32 acc)", idesc(1, 0, 0, 128, 256), 128, 256, 16, grid);
    run<0>("f16  M128 N64  K16 (fp32 acc)", idesc(1, 0, 0, 128, 64), 128, 64, 16, grid);
    run<1>("i8   M128 N128 K32 (s32 acc)", idesc(2, 1, 1, 128, 128), 128, 128, 32, grid);
    run<1>("i8   M128 N256 K32 (s32 acc)", idesc(2, 1, 1, 128, 256), 128, 256, 32, grid);
    run<1>("i8 u8xs8 M128 N128 K32", idesc(2, 0, 1, 128, 128), 128, 128, 32, grid);
    run<2>("e4m3 M128 N128 K32 (fp32 acc)", idesc(1, 0, 0, 128, 128), 128, 128, 32, grid);
    run<2>("e4m3 M128 N256 K32 (fp32 acc)", idesc(1, 0, 0, 128, 256), 128, 256, 32, grid);
    run<3>("tf32 M128 N128 K8  (fp32 acc)", idesc(1, 2, 2, 128, 128), 128, 128, 8, grid);
  }
  return 0;
}
