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

// Same MMA stream, but warp 2 keeps `depth` bulk copies (global -> shared, 16 KB each) in flight into a separate
// ring while the MMAs run, and warps 3.. write / read a scratch area with st/ld.shared: does shared-memory traffic
// from TMA fills and epilogue staging slow the tensor pipe's operand fetch (the in-kernel N=128 rate is 84 cycles)?
template <int N>
__global__ void __launch_bounds__(256, 1) umma_contended_kernel(int iters, const uint8_t* __restrict__ src, int depth,
                                                                int stagers, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, cbar[8];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int OPS = 128 * 128 + N * 128;
  uint8_t* ring = smem + ((OPS + 1023) / 1024) * 1024;          // [8][16 KB]
  uint8_t* scratch = ring + 8 * 16384;                          // [16 KB]
  for (int i = threadIdx.x; i < OPS / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int i = 0; i < 8; ++i) mbar_init(&cbar[i], 1);
    fence_barrier_init();
    stop = 0;
  }
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
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma<false>(tmem, ad + 2 * k, bd + 2 * k, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long dt = clock64() - t0;
    if (lane == 0) { stop = 1; if (blockIdx.x == 0) out[0] = dt; }
  } else if (warp == 2 && depth > 0) {
    // bulk-copy stream: keep `depth` 16 KB copies in flight
    uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int issued = 0;
    const uint8_t* g = src + (size_t)blockIdx.x * (1 << 20);
    while (!stop) {
      const int s = issued % depth;
      if (issued >= depth) { mbar_wait(&cbar[s], ph[s]); ph[s] ^= 1; }
      if (elect_one()) {
        mbar_expect_tx(&cbar[s], 16384);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(ring + s * 16384)),
                     "l"(g + (size_t)(issued & 63) * 16384), "r"(16384), "r"(smem_u32(&cbar[s]))
                     : "memory");
      }
      __syncwarp();
      ++issued;
    }
    for (int s = 0; s < depth && s < issued; ++s) mbar_wait(&cbar[s], ph[s]);      // drain before exit
  } else if (warp >= 3 && warp < 3 + stagers) {
    // epilogue-like staging traffic: 16-byte stores and loads on a scratch area
    uint32_t addr = smem_u32(scratch) + ((threadIdx.x * 16) & 16383);
    uint32_t acc = 0;
    while (!stop) {
      sts_b4(addr, acc, acc, acc, acc);
      const uint4 v = lds_u4(addr ^ 2048);
      acc += v.x;
    }
    if (acc == 0x12345678u) out[1] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N>
static void run_contended(int depth, int stagers) {
  long long* d; uint8_t* src;
  cudaMalloc(&d, 2 * sizeof(long long));
  cudaMalloc(&src, (size_t)148 << 20);
  cudaMemset(src, 0, (size_t)148 << 20);
  const int iters = 2048;
  const int smem = ((128 * 128 + N * 128 + 1023) / 1024) * 1024 + 9 * 16384;
  cudaFuncSetAttribute(umma_contended_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  umma_contended_kernel<N><<<148, 256, smem>>>(iters, src, depth, stagers, d);
  long long h = 0;
  cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const cudaError_t e = cudaGetLastError();
  printf("N=%3d  bulk copies in flight=%d  staging warps=%d: %.1f cycles per MMA (ideal %d)  %s\n", N, depth, stagers,
         (double)h / (iters * 4), N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d); cudaFree(src);
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
  // shared-memory contention: TMA-like fills and epilogue-like staging next to the MMAs
  run_contended<128>(0, 0); run_contended<128>(4, 0); run_contended<128>(8, 0); run_contended<128>(0, 4); run_contended<128>(8, 4);
  run_contended<64>(0, 0); run_contended<64>(8, 4);
  run_contended<256>(0, 0); run_contended<256>(8, 4);
  return 0;
}
