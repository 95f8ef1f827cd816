// fp64 tensor-core (DMMA) issue rate per shape on one SM-full of warps:  nvcc -gencode arch=compute_100a,code=sm_100a
//   -O3 -o dmma_rate dmma_rate.cu && ./dmma_rate
// Each warp runs ITER rounds of NACC independent mma.sync of the given shape; prints MAC/clk/SM for 4/8/16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2000, NACC = 8;

template <int SHAPE> __device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
  if constexpr (SHAPE == 0) {         // m8n8k4
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
  } else if constexpr (SHAPE == 1) {  // m16n8k4
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
  } else if constexpr (SHAPE == 2) {  // m16n8k8
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
  } else {                            // m16n8k16
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, "
                 "{%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]),
                   "d"(b[1]), "d"(b[2]), "d"(b[3]));
  }
}

template <int SHAPE> __global__ void k(double* out, long long* clk) {
  double a[8], b[4], c[NACC][4];
  for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
  for (int i = 0; i < 4; ++i) b[i] = 0.5 + threadIdx.x * 1e-9 + i;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma<SHAPE>(c[j], a, b);
  }
  __syncthreads();
  const long long t1 = clock64();
  double s = 0;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int SHAPE> void run(const char* name, int macs) {
  double* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  for (int warps : {4, 8, 16, 32}) {
    k<SHAPE><<<148, warps * 32>>>(out, clk);
    k<SHAPE><<<148, warps * 32>>>(out, clk);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double total_macs = double(ITER) * NACC * warps * macs;
    printf("%-10s %2d warps/SM: %8.0f clk -> %7.1f MAC/clk/SM = %5.1f TFLOP/s at 148 SMs x 1.9 GHz, %5.2f clk per MMA per warp\n", name,
           warps, avg, total_macs / avg, total_macs / avg * 2 * 148 * 1.9e9 / 1e12, avg / (double(ITER) * NACC));
  }
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<0>("m8n8k4", 8 * 8 * 4);
  run<1>("m16n8k4", 16 * 8 * 4);
  run<2>("m16n8k8", 16 * 8 * 8);
  run<3>("m16n8k16", 16 * 8 * 16);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
