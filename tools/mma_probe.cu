// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, M = 128, K = 16, bf16) as a function of where the A operand
// lives (shared memory = "SS", tensor memory = "TS"), of N, and of concurrent shared-memory traffic from other warps.
// Answers the question behind K4 / K9: is a 128 x 128 SS-mode MMA bound by the tensor pipe (64 cycles) or by the
// shared-memory reads of its operands (8 KB per MMA)?   Build / run: tools/mma_probe.sh (results: profiles/r2_mma_probe.md)
#include <cstdio>
#include <cstdlib>

#include "../mmmm_b200/csrc/common.cuh"

using namespace vex;

constexpr int TILE = 128 * 128 * 2;  // 32 KB
constexpr int ATOM = 128 * 64 * 2;
constexpr int NSLOT = 5;             // A tile + 4 B tiles
constexpr int SMEM = NSLOT * TILE + 2048;

// mode: 0 SS (A K-major, B K-major: S = Q K^T), 1 SS (B MN-major: O += P V), 2 TS (A in TMEM, B MN-major),
//       3 TS (A in TMEM, B K-major)
// traffic: 0 none, 1 = 8 warps storing 16-byte vectors into the spare tile back to back, 2 = 8 warps loading
template <int N>
__global__ void __launch_bounds__(384, 1) probe(int mode, int traffic, int n_blocks, long long* cycles_out, int chains) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NSLOT * TILE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  volatile int* stop = reinterpret_cast<volatile int*>(bar + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NSLOT * TILE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    *stop = 0;
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1 || (warp == 3 && chains == -2)) {
    const int second = warp == 3;
    if (chains == -2 && !second) __nanosleep(0);
    if (elect_one_sync()) {
      const uint32_t idesc = umma_idesc_bf16(128, N, 0, (mode == 1 || mode == 2) ? 1 : 0);
      const uint64_t dA = umma_desc_kmajor_sw128(smem_u32(smem));
      const uint64_t dBk = umma_desc_kmajor_sw128(smem_u32(smem + TILE));
      const uint64_t dBm = umma_desc_mnmajor_sw128(smem_u32(smem + TILE), ATOM, 1024);
      const long long t0 = clock64();
      for (int b = 0; b < n_blocks; ++b) {
        const uint64_t slot = static_cast<uint64_t>((b % 3) * (TILE >> 4));
        // chains == 1: the 8 K-steps of a block accumulate into one tile (a dependent chain, like one S or PV GEMM);
        // chains == 2 / 4: consecutive MMAs go to different accumulators (round-robin), each accumulator still receives
        // every chains-th K-step -- the interleaving two issuer warps (or one issuer walking two tiles) would produce
        const uint32_t tD0 = tmem + (chains == -2 ? second * 256 : (b & 1) * 128);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t tD = chains <= 1 ? tD0 : tmem + (kk % chains) * (N > 128 ? 128 : N);
          const uint64_t offk = static_cast<uint64_t>(((kk >> 2) * ATOM + (kk & 3) * 32) >> 4);
          const uint64_t offm = static_cast<uint64_t>(kk * (2048 >> 4));
          if (mode == 0) umma_ss(tD, dA + offk, dBk + slot + offk, idesc, kk > 0);
          else if (mode == 1) umma_ss(tD, dA + offk, dBm + slot + offm, idesc, kk > 0);
          else if (mode == 2) umma_ts(tD, tmem + 128 + kk * 8, dBm + slot + offm, idesc, kk > 0);
          else umma_ts(tD, tmem + 128 + kk * 8, dBk + slot + offk, idesc, kk > 0);
        }
      }
      umma_commit(bar + second);
      mbar_wait(bar + second, 0);
      const long long t1 = clock64();
      if (!second) cycles_out[blockIdx.x] = t1 - t0;
      if (!second) *stop = 1;
    }
    __syncwarp();
  } else if (warp >= 4 && traffic != 0) {
    // spare tile (slot 4): every lane its own 16-byte chunk, swizzled like the P stores of the attention kernel
    const uint32_t base = smem_u32(smem + 4 * TILE) + (warp - 4) * 4096 + lane * 128;
    uint32_t acc = 0;
    long long n = 0;
    while (!*stop) {
#pragma unroll
      for (int c16 = 0; c16 < 8; ++c16) {
        const uint32_t addr = base + ((c16 ^ (lane & 7)) << 4);
        if (traffic == 1) {
          asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(addr), "r"(acc) : "memory");
        } else {
          uint32_t a, b2, c, d;
          asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b2), "=r"(c), "=r"(d) : "r"(addr));
          acc += a ^ b2 ^ c ^ d;
        }
      }
      ++n;
    }
    if (acc == 0x12345678u) cycles_out[gridDim.x + blockIdx.x] = n;  // keep the loads alive
    if (lane == 0 && warp == 4) cycles_out[gridDim.x + blockIdx.x] = n;  // iterations of 8 x 512 B per warp
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N>
void run(const char* name, int mode, int traffic, int grid, int chains = 1) {
  long long* d;
  cudaMalloc(&d, 2 * grid * sizeof(long long));
  cudaMemset(d, 0, 2 * grid * sizeof(long long));
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int n_blocks = 512;
  for (int rep = 0; rep < 2; ++rep) probe<N><<<grid, 384, SMEM>>>(mode, traffic, n_blocks, d, chains);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: %s\n", name, cudaGetErrorString(e));
    exit(1);
  }
  long long* h = static_cast<long long*>(malloc(2 * grid * sizeof(long long)));
  cudaMemcpy(h, d, 2 * grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double s = 0, it = 0;
  for (int i = 0; i < grid; ++i) s += h[i], it += h[grid + i];
  const double cyc = s / grid / (n_blocks * 8.0 * (chains == -2 ? 2 : 1));
  // traffic bytes per cycle per SM: 8 warps x iterations x 8 x 512 B
  const double tb = traffic ? (it / grid) * 8.0 * 8 * 512 / (s / grid) : 0.0;
  printf("| %-34s | N=%3d | chains %d | traffic %d | %7.1f cyc/MMA | ideal %3d | side traffic %6.1f B/clk |\n", name, N, chains, traffic, cyc,
         128 * N / 256, tb);
  cudaFree(d);
  free(h);
}

int main(int argc, char** argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 148;
  printf("two issuer warps, one dependent 8-MMA chain each at a time (cycles per MMA over both)\n");
  run<128>("SS  A smem K-major, B K-major", 0, 0, grid, -2);
  run<128>("SS  A smem K-major, B MN-major", 1, 0, grid, -2);
  run<128>("TS  A tmem, B MN-major", 2, 0, grid, -2);
  run<64>("SS  A smem K-major, B K-major", 0, 0, grid, -2);
  printf("one issuer, consecutive MMAs alternate between accumulators\n");
  for (int chains = 2; chains <= 2; chains += 2) {
    run<128>("SS  A smem K-major, B K-major", 0, 0, grid, chains);
    run<128>("SS  A smem K-major, B MN-major", 1, 0, grid, chains);
    run<128>("TS  A tmem, B MN-major", 2, 0, grid, chains);
    run<128>("TS  A tmem, B K-major", 3, 0, grid, chains);
    run<64>("SS  A smem K-major, B K-major", 0, 0, grid, chains);
    run<64>("TS  A tmem, B K-major", 3, 0, grid, chains);
    run<128>("SS  A smem K-major, B K-major", 0, 1, grid, chains);
    run<128>("TS  A tmem, B MN-major", 2, 1, grid, chains);
  }
  for (int traffic = 0; traffic < 2; ++traffic) {
    run<128>("SS  A smem K-major, B K-major", 0, traffic, grid);
    run<128>("SS  A smem K-major, B MN-major", 1, traffic, grid);
    run<128>("TS  A tmem, B MN-major", 2, traffic, grid);
    run<128>("TS  A tmem, B K-major", 3, traffic, grid);
    run<64>("SS  A smem K-major, B K-major", 0, traffic, grid);
    run<64>("TS  A tmem, B K-major", 3, traffic, grid);
    run<256>("SS  A smem K-major, B K-major", 0, traffic, grid);
    run<256>("TS  A tmem, B K-major", 3, traffic, grid);
  }
  return 0;
}
