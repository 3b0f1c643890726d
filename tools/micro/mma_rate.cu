// Micro-benchmark: cycles per tcgen05.mma (M128 x N x K16, bf16) for SS / TS operand modes and K- / MN-major B,
// one CTA per SM, operands resident in shared memory (contents irrelevant).  Build + run: tools/micro/run_mma_rate.sh
#include <cstdio>
#include <cuda.h>
#include "../../mmmm_b200/csrc/common.cuh"
using namespace vex;

__device__ __forceinline__ void umma_ts_(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b),
               "r"(idesc), "r"(acc) : "memory");
}

template <int MODE, int N>  // 0: SS, B K-major; 1: SS, B MN-major; 2: TS, B MN-major; 3: TS, B K-major
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int iters, int sts_traffic) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tbase, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tbase;
  constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, (MODE == 1 || MODE == 2) ? 1 : 0);
  const uint32_t aA = smem_u32(smem), aB = smem_u32(smem + 32768);
  if (warp == 0) {
    const uint64_t dA = umma_desc_kmajor_sw128(aA);
    const uint64_t dBk = umma_desc_kmajor_sw128(aB), dBm = umma_desc_mnmajor_sw128(aB, 16384, 1024);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t offk = static_cast<uint64_t>(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
          const uint64_t offm = static_cast<uint64_t>(kk * (2048 >> 4));
          if (MODE == 0) umma_ss(tmem, dA + offk, dBk + offk, idesc, 1);
          if (MODE == 1) umma_ss(tmem, dA + offk, dBm + offm, idesc, 1);
          if (MODE == 2) umma_ts_(tmem, tmem + 256 + kk * 8, dBm + offm, idesc, 1);
          if (MODE == 3) umma_ts_(tmem, tmem + 256 + kk * 8, dBk + offk, idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (sts_traffic < 0) {
    // competing TMEM traffic from the other three warps (like the softmax warps pulling S / rescaling O):
    // -1: tcgen05.ld of columns [256, 384) in a loop; -2: ld + st of the same columns
    const uint32_t t = tmem + (static_cast<uint32_t>(warp * 32) << 16) + 256;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
      uint32_t r[32];
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        tmem_ld_32x32b_x32(t + c * 32, r);
        tmem_ld_wait();
        acc += r[0] + r[31];
        if (sts_traffic == -2) {
          tmem_st_32x32b_x32(t + c * 32, r);
          tmem_st_wait();
        }
      }
    }
    if (acc == 0x12345678u) out[1] = acc;
  } else if (sts_traffic) {
    // competing generic-proxy shared-memory stores (like the P tile writes): 16 B per lane per store
    uint32_t addr = smem_u32(smem + 65536) + threadIdx.x * 16;
    for (int it = 0; it < iters * sts_traffic; ++it) {
#pragma unroll
      for (int k = 0; k < 16; ++k)
        asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr + k * 2048), "r"(it) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE, int N>
void run(const char* name, long long* d, int sts) {
  const int iters = 2000;
  cudaFuncSetAttribute(mma_rate<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) mma_rate<MODE, N><<<148, 128, 200 * 1024>>>(d, iters, sts);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d sts=%d: %7.1f cycles per MMA (floor %d)  %s\n", name, N, sts, double(h) / (iters * 8.0), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  run<0, 128>("SS  A K-major, B K-major", d, 0);
  run<1, 128>("SS  A K-major, B MN-major", d, 0);
  run<2, 128>("TS  A tmem,    B MN-major", d, 0);
  run<3, 128>("TS  A tmem,    B K-major", d, 0);
  run<0, 256>("SS  A K-major, B K-major", d, 0);
  run<2, 256>("TS  A tmem,    B MN-major", d, 0);
  run<0, 128>("SS  K/K + st.shared traffic", d, 1);
  run<2, 128>("TS  MN  + st.shared traffic", d, 1);
  run<0, 128>("SS  K/K + 4x st.shared traffic", d, 4);
  run<0, 128>("SS  K/K + 3 warps tcgen05.ld", d, -1);
  run<2, 128>("TS  MN  + 3 warps tcgen05.ld", d, -1);
  run<0, 128>("SS  K/K + 3 warps tcgen05.ld+st", d, -2);
  return 0;
}
