// Outbound bandwidth of one SM, shared memory -> global (L2-resident ring), every SM busy:
//   (a) cp.async.bulk.global.shared::cta  (TMA engine), 1 or 3 issuing threads, 8 / 24 KB per copy, K copies in flight
//   (b) st.global.v4 from 8 warps (LSU path) + one gpu-scope fence per 24 KB by a helper thread
//   (c) inbound: cp.async.bulk.shared.global (24 KB per copy, mbarrier completion), K in flight
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o smem_out_rate smem_out_rate.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int NTHR, int BYTES, int INFLIGHT>
__global__ void tma_store_kernel(uint8_t* ring, int iters, long long* clk) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  uint8_t* dst = ring + static_cast<size_t>(blockIdx.x) * (32 * 24576);
  const long long t0 = clock64();
  if (threadIdx.x < NTHR) {
    for (int it = 0; it < iters; ++it) {
      const int slot = it & 31;
      uint8_t* d = dst + slot * 24576 + threadIdx.x * BYTES;
      const uint32_t s = smem_u32(sm) + (it & 3) * 24576 + threadIdx.x * BYTES;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(s), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group %0;" ::"n"(INFLIGHT) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

__global__ void stg_store_kernel(uint8_t* ring, int iters, int fence, long long* clk, unsigned* flags) {
  extern __shared__ __align__(1024) uint8_t sm[];
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  __syncthreads();
  uint8_t* dst = ring + static_cast<size_t>(blockIdx.x) * (32 * 24576);
  const long long t0 = clock64();
  // 256 threads: each copies 24576 / 256 = 96 B = 6 x 16 B per iteration
  for (int it = 0; it < iters; ++it) {
    const int slot = it & 31;
    const uint4* s = reinterpret_cast<const uint4*>(sm + (it & 3) * 24576);
    uint4* d = reinterpret_cast<uint4*>(dst + slot * 24576);
#pragma unroll
    for (int k = 0; k < 6; ++k) d[k * 256 + threadIdx.x] = s[k * 256 + threadIdx.x];
    if (fence) {
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flags + blockIdx.x), "r"(it + 1) : "memory");
      }
    }
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

template <int INFLIGHT>
__global__ void tma_load_kernel(const uint8_t* ring, int iters, long long* clk) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) unsigned long long bars[INFLIGHT];
  if (threadIdx.x == 0) {
    for (int i = 0; i < INFLIGHT; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // every CTA reads the ring of CTA (blockIdx.x / 8) * 8  -> 8 readers per ring, like a tile column
  const uint8_t* src = ring + static_cast<size_t>((blockIdx.x / 8) * 8) * (32 * 24576);
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int it = 0; it < iters + INFLIGHT; ++it) {
      const int b = it % INFLIGHT;
      if (it >= INFLIGHT) {
        const uint32_t par = ((it / INFLIGHT) - 1) & 1;
        uint32_t done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                       : "=r"(done) : "r"(smem_u32(&bars[b])), "r"(par) : "memory");
      }
      if (it < iters) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[b])), "r"(24576) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm) + b * 24576), "l"(src + (it & 31) * 24576), "r"(24576), "r"(smem_u32(&bars[b])) : "memory");
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}

static double report(const char* name, long long* dclk, int nblk, int iters, double bytes_per_iter) {
  long long h[148];
  cudaMemcpy(h, dclk, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; double avg = 0;
  for (int i = 0; i < nblk; ++i) { mx = h[i] > mx ? h[i] : mx; avg += h[i]; }
  avg /= nblk;
  printf("%-58s  %7.0f clk/iter  %6.1f B/clk/SM (avg)  %6.1f (slowest SM)\n", name, avg / iters, bytes_per_iter * iters / avg,
         bytes_per_iter * iters / mx);
  return avg;
}

int main() {
  const int nblk = 148, iters = 2000;
  uint8_t* ring; long long* clk; unsigned* flags;
  cudaMalloc(&ring, static_cast<size_t>(nblk) * 32 * 24576);
  cudaMalloc(&clk, nblk * sizeof(long long));
  cudaMalloc(&flags, nblk * sizeof(unsigned));
  cudaMemset(ring, 0, static_cast<size_t>(nblk) * 32 * 24576);
  const size_t smem = 128 * 1024;
#define RUN_TMA(NTHR, BYTES, INF, label)                                                                              \
  cudaFuncSetAttribute(tma_store_kernel<NTHR, BYTES, INF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
  tma_store_kernel<NTHR, BYTES, INF><<<nblk, 128, smem>>>(ring, iters, clk);                                         \
  cudaDeviceSynchronize();                                                                                            \
  report(label, clk, nblk, iters, double(NTHR) * BYTES);
  RUN_TMA(1, 24576, 0, "TMA store 1 thread x 24 KB, 1 in flight");
  RUN_TMA(1, 24576, 2, "TMA store 1 thread x 24 KB, 3 in flight");
  RUN_TMA(1, 24576, 6, "TMA store 1 thread x 24 KB, 7 in flight");
  RUN_TMA(3, 8192, 2, "TMA store 3 threads x 8 KB, 3 in flight each");
  RUN_TMA(1, 8192, 6, "TMA store 1 thread x 8 KB, 7 in flight");
  RUN_TMA(1, 2048, 6, "TMA store 1 thread x 2 KB, 7 in flight");
  RUN_TMA(12, 2048, 2, "TMA store 12 threads x 2 KB, 3 in flight each");
  cudaFuncSetAttribute(stg_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  stg_store_kernel<<<nblk, 256, smem>>>(ring, iters, 0, clk, flags);
  cudaDeviceSynchronize();
  report("STG.128 x 256 threads, 24 KB per iter, no fence", clk, nblk, iters, 24576.0);
  stg_store_kernel<<<nblk, 256, smem>>>(ring, iters, 1, clk, flags);
  cudaDeviceSynchronize();
  report("STG.128 x 256 threads, 24 KB per iter, bar + fence + flag", clk, nblk, iters, 24576.0);
  cudaFuncSetAttribute(tma_load_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tma_load_kernel<2><<<nblk, 128, smem>>>(ring, iters, clk);
  cudaDeviceSynchronize();
  report("TMA load 24 KB, 2 in flight, 8 readers per ring", clk, nblk, iters, 24576.0);
  cudaFuncSetAttribute(tma_load_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tma_load_kernel<4><<<nblk, 128, smem>>>(ring, iters, clk);
  cudaDeviceSynchronize();
  report("TMA load 24 KB, 4 in flight, 8 readers per ring", clk, nblk, iters, 24576.0);
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
