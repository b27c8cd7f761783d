// Micro-benchmark (tuning aid, not part of libdt_b200.so): cycles per tcgen05.mma (cta_group::1, M=128, K=16 bf16)
// as a function of N and of how many MMAs are issued back to back, to decide between 256-row M tiles and
// cta_group::2 pairs for the narrow layers (NOTES.md "Queue" item 1).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I detectandtrack_b200/csrc experiments/umma_rate.cu -o /tmp/umma_rate
//   /tmp/umma_rate            (prints one line per (N, accumulators))
#include <cstdio>
#include <cuda_bf16.h>
#include "tc_common.cuh"

using namespace dt::tc;

// One CTA per SM.  Operands: A 128 x 64 bf16 (one 128-byte-swizzled k-block), B N x 64 bf16, both zero.
// `nacc` accumulators (each N columns) are cycled so that back-to-back MMAs either chain on one accumulator
// (nacc = 1, the conv kernel's situation) or are independent (nacc = 2: the 256-row-tile proposal).
template <int N>
__global__ void __launch_bounds__(64, 1) umma_rate_kernel(int iters, int nacc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (128 * 128 + N * 128) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(128, N, 1);
    const uint32_t a = smem_u32(smem), b = a + 128 * 128;
    const uint64_t ad = make_sw128_kmajor_desc(a), bd = make_sw128_kmajor_desc(b);
    uint32_t phase = 0;
    long long best = 1ll << 60;
    for (int rep = 0; rep < 5; ++rep) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < iters; ++i) {
          const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma<false>(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
        }
        umma_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, phase);
      phase ^= 1;
      const long long dt = clock64() - t0;
      best = dt < best ? dt : best;
    }
    if (lane == 0 && blockIdx.x == 0) out[0] = best;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N>
static void run(int nacc) {
  long long* d;
  cudaMalloc(&d, sizeof(long long));
  const int iters = 256;                           // x 4 MMAs each
  const int smem = 128 * 128 + N * 128;
  cudaFuncSetAttribute(umma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_rate_kernel<N><<<148, 64, smem>>>(iters, nacc, d);
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const cudaError_t e = cudaGetLastError();
  printf("N=%3d accumulators=%d: %.1f cycles per MMA (ideal %d)  %s\n", N, nacc, (double)h / (iters * 4), N / 2,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64>(1); run<64>(2); run<64>(4);
  run<128>(1); run<128>(2);
  run<256>(1); run<256>(2);
  return 0;
}
