// Conv3d / Conv2d / FC as ONE implicit-GEMM kernel on the 5th-gen tensor cores (tcgen05),
// NDHWC activations, fused AffineChannel (+bias) + residual / top-down-upsample add + ReLU.
//
// Replaces, on the reference's hot path:
//   Caffe2 Conv / ConvNd engine=CUDNN        (lib/modeling/detector.py:318-322,410-436 and every
//                                             call site listed in SURVEY.md §2.3 K5)
//   AffineChannelNdOp<float,CUDAContext>     (lib/ops/affine_channel_nd_op.cu:19-70)  -> epilogue
//   Relu / Sum (residual)                    (lib/modeling/ResNet3D.py:37,47,147-152) -> epilogue
//   UpsampleNearest + Sum (FPN top-down)     (lib/modeling/FPN3D.py:211-222)          -> epilogue
//   FC (cuBLAS)                              (lib/modeling/head_builder.py:33-36)     -> kT=kH=kW=1
//
// GEMM view: D[m, n] = sum_{tap, c} A[m @ tap, c] * W[tap, n, c]
//   m = output position (img, t, ho, wo) tiled as TH x TW spatial patches (TH*TW <= 128 rows)
//   n = output channel, BLOCK_N in {32, 64, 128, 256};  k-block = one tap x 128 bytes of channels
//   A tile: one 5-D TMA box (C=128B, TW, TH, 1, 1) at the tap-shifted coordinate; out-of-bounds
//           elements (spatial / temporal zero padding, ragged edge tiles, channel tail) are
//           zero-filled by the TMA unit, so padding costs no instructions and no branches.
//   W tile: 3-D TMA box (C=128B, BLOCK_N, 1) of the pre-packed [tap][Cout][Cin] weights.
//   Both land in the 128B-swizzled K-major layout tcgen05.mma reads directly.
// Roles (576 threads, 1 CTA / SM, persistent over tiles):
//   warp 0   : TMA producer (one elected lane)          smem ring: full[]/empty[] mbarriers
//   warp 1   : TMEM allocator + MMA issuer (one lane)   tcgen05.mma -> TMEM, tcgen05.commit
//   warps 2-17: epilogue; TMEM -> registers (tcgen05.ld), scale/bias/residual/ReLU in fp32, staged in
//              128B-swizzled smem chunks and written with TMA stores.  Two TMEM accumulator stages
//              overlap it with the next tile's MMAs.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/dt_b200.h"
#include <cuda_bf16.h>

namespace dt {

using namespace tc;

struct ConvKernelParams {
  // output geometry
  int N, To, Ho, Wo, Cout;
  // filter
  int kT, kH, kW, sT, sH, sW, pT, pH, pW;
  int kchunks;                 // ceil(Cin / BK)
  // tiling
  int TH, TW, TT, TB;          // rows of one M tile = TB images x TT frames x TH x TW positions (<= 128)
  int tiles_h, tiles_w, tiles_t, tiles_b, tiles_n, total_tiles;
  uint32_t a_bytes;            // TB*TT*TH*TW*128
  // epilogue
  const float* scale;          // [Cout] or null (1)
  const float* bias;           // [Cout] or null (0)
  const void* residual;        // same dtype as out, or null
  int res_mode;                // 0 none, 1 same shape, 2 nearest-2x upsample of (Ho/2, Wo/2)
  int res_ld;
  int relu;
  void* out;
  int out_ld;                  // elements between consecutive positions
  int out_f32;                 // 1: fp32 output, 0: bf16
  // 3xTF32 ("fp32-accurate") mode: activations / weights are stored as [hi | lo] tf32 pairs along the
  // channel axis; D = A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (the lo*lo term is below fp32 resolution).
  int split_in;                // 1: three k-blocks per (tap, channel chunk) with the offsets below
  int a_lo_off, b_lo_off;      // element offsets of the lo halves in the x rows / packed weight rows
  int split_out;               // 1: write hi at channel c and lo at out_lo_off + c (fp32)
  int out_lo_off, res_lo_off;  // lo-half offsets of the output / residual rows
  int nstages, ncbuf;          // smem split chosen per layer: operand ring depth / output staging buffers
  int row_planes;              // conv1: input rows de-interleaved by parity, filter row kh -> plane kh & 1, row + kh >> 1
  int nrbuf;                   // > 0: bf16 residual chunks arrive by TMA in a ring of this many staged chunks
  int cgroup;                  // output chunks staged per named-barrier pair / TMA commit group (divides ncbuf)
  int round_tf32;              // fp32 output is rounded (RNE) to tf32 so the next tcgen05 kind::tf32 MMA,
                               // which TRUNCATES its 32-bit operands, sees exactly representable values
};

__device__ __forceinline__ float round_to_tf32(float v) {
  uint32_t u = __float_as_uint(v);
  u += 0xFFFu + ((u >> 13) & 1u);
  return __uint_as_float(u & 0xFFFFE000u);
}

constexpr int EPI_WARPS = 16;                         // 4 TMEM lane groups x 4 column quarters
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int CONV_THREADS = 64 + EPI_THREADS;        // + TMA producer warp + MMA issuer warp

__device__ __forceinline__ void epi_bar_sync() {      // named barrier 1: the epilogue warps only
  asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
}

template <int BN>
struct ConvCfg {
  static constexpr int A_BYTES = 128 * 128;            // 128 rows x 128 B
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int MAX_STAGES = 8;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int C_BYTES = 128 * 128;            // one staged output chunk: 128 rows x 128 B
  static constexpr int BAR_BYTES = 384;                // mbarriers + TMEM base pointer
  static constexpr int FIXED_BYTES = BAR_BYTES + 2 * BN * 4 /*scale, bias*/;
  static constexpr int BUDGET = 227 * 1024;
  // K-heavy layers want a deep operand ring; K-light (HBM-bound) layers want output staging buffers so the
  // epilogue never waits for a TMA store to drain, and (with a residual) a ring of prefetched residual chunks.
  static void split(int kiters, bool res_tma, int* stages, int* ncbuf, int* nrbuf) {
    int c = (kiters >= 12) ? ((BN >= 256) ? 1 : 2) : 4;
    const int r = res_tma ? ((kiters >= 12) ? 2 : 4) : 0;
    int st = (BUDGET - FIXED_BYTES - (c + r) * C_BYTES) / STAGE_BYTES;
    if (st > MAX_STAGES) st = MAX_STAGES;
    *stages = st; *ncbuf = c; *nrbuf = r;
  }
  static int smem_bytes(int stages, int ncbuf, int nrbuf) {
    return stages * STAGE_BYTES + (ncbuf + nrbuf) * C_BYTES + FIXED_BYTES;
  }
};

template <int BN, bool TF32>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const ConvKernelParams p) {
  using Cfg = ConvCfg<BN>;
  const int STAGES = p.nstages;
  constexpr int BK = TF32 ? 32 : 64;                 // elements per 128-byte k-block
  // SWIZZLE_128B operands need 1024-byte aligned tiles: the dynamic window is declared with that alignment
  // (no static shared memory in this kernel) and checked once below instead of spending a kilobyte on slack
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* cbuf = smem + STAGES * Cfg::STAGE_BYTES;                  // [NCBUF][128 rows][128 B], 128B-swizzled
  uint8_t* rbuf = cbuf + p.ncbuf * Cfg::C_BYTES;                     // [NRBUF] residual chunks, same layout
  uint64_t* bars = reinterpret_cast<uint64_t*>(rbuf + p.nrbuf * Cfg::C_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + STAGES;             // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint64_t* r_full = bars + 2 * STAGES + 4;    // [4]
  uint64_t* r_empty = bars + 2 * STAGES + 8;   // [4]
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 12);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::BAR_BYTES);   // [BN]
  float* s_bias = s_scale + BN;                                                                      // [BN]
  if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    prefetch_tmap(&tmC);
    if (p.nrbuf > 0) prefetch_tmap(&tmR);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], EPI_WARPS); }
    for (int s = 0; s < 4; ++s) { mbar_init(&r_full[s], 1); mbar_init(&r_empty[s], EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_base_smem);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_base_smem;

  const int taps = p.kT * p.kH * p.kW;
  const int kiters = taps * p.kchunks * (p.split_in ? 3 : 1);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop (warp-uniform control flow and addresses, so descriptors / coordinates
    // stay in uniform registers); one elected lane issues.  A divergent `if (lane == 0)` around the loop
    // makes the compiler wrap every UTMALDG / UTCHMMA in an elect-broadcast "waterfall" loop, which costs
    // more than the MMAs of a narrow tile take to execute.
    int stage = 0; uint32_t phase = 0;
    int rslot = 0; uint32_t rphase = 0;
    const uint32_t smem_u = smem_u32(smem);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n;
      int mt = tile / p.tiles_n;
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int thi = mt % p.tiles_h; mt /= p.tiles_h;
      const int tti = mt % p.tiles_t;
      const int n = (mt / p.tiles_t) * p.TB;
      const int w_base = twi * p.TW * p.sW - p.pW;
      const int h_base = thi * p.TH * p.sH - p.pH;
      const int t_base = tti * p.TT * p.sT - p.pT;
      const int nsub = p.split_in ? 3 : 1;
      int tap = 0;
      for (int kt = 0; kt < p.kT; ++kt)
        for (int kh = 0; kh < p.kH; ++kh)
          for (int kw = 0; kw < p.kW; ++kw, ++tap)
            for (int kc = 0; kc < p.kchunks; ++kc)
              for (int sub = 0; sub < nsub; ++sub) {
                // sub 0: A_hi x B_hi, 1: A_lo x B_hi, 2: A_hi x B_lo
                const int ca = kc * BK + (sub == 1 ? p.a_lo_off : 0);
                const int cb = kc * BK + (sub == 2 ? p.b_lo_off : 0);
                mbar_wait(&empty[stage], phase ^ 1);
                if (elect_one()) {
                  const uint32_t a_dst = smem_u + stage * Cfg::STAGE_BYTES;
                  const uint32_t bar = smem_u32(&full[stage]);
                  mbar_expect_tx_u(bar, p.a_bytes + Cfg::B_BYTES);
                  if (p.row_planes) tma_load_5d_u(a_dst, &tmA, bar, ca, w_base, h_base + (kh >> 1), kh & 1, n);
                  else tma_load_5d_u(a_dst, &tmA, bar, ca, w_base + kw, h_base + kh, t_base + kt, n);
                  tma_load_3d_u(a_dst + Cfg::A_BYTES, &tmB, bar, cb, nt * BN, tap);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
              }
      if (p.nrbuf > 0) {
        // residual chunks of this tile (bf16, same box as the output chunks), consumed in order by the epilogue
        const int nbase = nt * BN;
        const int ncols = min(BN, p.Cout - nbase);
        for (int cc = 0; cc < ncols; cc += 64) {
          mbar_wait(&r_empty[rslot], rphase ^ 1);
          if (elect_one()) {
            const uint32_t bar = smem_u32(&r_full[rslot]);
            mbar_expect_tx_u(bar, p.a_bytes);
            tma_load_5d_u(smem_u32(rbuf) + rslot * Cfg::C_BYTES, &tmR, bar, nbase + cc, twi * p.TW, thi * p.TH,
                          tti * p.TT, n);
          }
          __syncwarp();
          if (++rslot == p.nrbuf) { rslot = 0; rphase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp, one elected lane issues) =====================
    constexpr uint32_t idesc = make_idesc(128, BN, TF32 ? 2 : 1);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_u + as * BN;
      for (int ki = 0; ki < kiters; ++ki) {
        mbar_wait(&full[stage], phase);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u + stage * Cfg::STAGE_BYTES;
        const uint64_t adesc = make_sw128_kmajor_desc(a_addr);
        const uint64_t bdesc = make_sw128_kmajor_desc(a_addr + Cfg::A_BYTES);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)          // 4 x 32 B = one 128-byte swizzle row of K
            umma<TF32>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (ki | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);          // frees the smem slot when these MMAs retire
          if (ki == kiters - 1) umma_commit(&tmem_full[as]);   // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..17) =====================
    // TMEM -> registers -> fp32 epilogue -> 128B-swizzled smem chunk -> one TMA store per chunk.
    // The loop is latency-bound per warp (dependent address / convert / store chains, named barriers), so
    // it runs on SIXTEEN warps (four per scheduler): warp w reads TMEM lane group (w & 3) and owns column
    // quarter ((w - 2) >> 2) of every staged 128-byte row (32 bytes: 16 bf16 or 8 fp32 outputs).
    // The TMA store writes whole 128-byte lines and clips rows / channels outside the tensor, so ragged
    // tiles need no predication on the store side.  All layer constants live in registers, ring positions
    // are counters (no divisions), scale/bias come from shared memory with explicit ld.shared.v4, ReLU rides
    // on the bf16 pack (cvt.rn.relu), and the TMEM load of chunk c+1 is issued as soon as chunk c's
    // accumulators have been consumed.
    const int lg = warp & 3;                   // TMEM lane group this warp may access
    const int part = (warp - 2) >> 2;          // which 32-byte quarter of the staged row this warp fills
    const int row = lg * 32 + lane;            // accumulator row == TMEM lane == staging row
    int rr = row;
    const int tw = rr % p.TW; rr /= p.TW;
    const int th = rr % p.TH; rr /= p.TH;
    const int tl = rr % p.TT;
    const int nl = rr / p.TT;                  // >= TB for the unused tail rows of a short tile
    const bool store_warp = (warp == 2);       // one elected lane of warp 2 issues / tracks the TMA stores
    const int ep_tid = threadIdx.x - 64;       // 0..511
    const bool out_f32 = p.out_f32 != 0;
    const bool split_out = p.split_out != 0;
    const bool relu = p.relu != 0;
    const int res_mode = p.res_mode;
    const int Cout = p.Cout;
    const int CW = out_f32 ? 32 : 64;          // output columns per 128-byte staged row
    const int colw = out_f32 ? 8 : 16;         // columns this warp owns per chunk
    const int ncbuf = p.ncbuf, cgroup = p.cgroup;
    const int bstep = split_out ? 2 : 1;       // split output: buffers (2i, 2i+1) hold the hi / lo chunk
    const int inflight = split_out ? 1 : ncbuf / cgroup - 1;   // commit groups that may stay pending
    const uint32_t cbuf_u32 = smem_u32(cbuf);
    const uint32_t rbuf_u32 = smem_u32(rbuf);
    const bool res_tma = p.nrbuf > 0;
    int rslot = 0; uint32_t rphase = 0;
    const uint32_t row_smem = (uint32_t)row * 128u;
    const uint32_t swz = (uint32_t)(row & 7);
    const uint32_t q0 = (((uint32_t)(2 * part)) ^ swz) << 4, q1 = (((uint32_t)(2 * part + 1)) ^ swz) << 4;
    const uint32_t sc_u32 = smem_u32(s_scale) + (uint32_t)(colw * part) * 4u;
    const uint32_t bi_u32 = sc_u32 + BN * 4u;
    int as = 0; uint32_t aphase = 0;
    int buf_idx = 0, gi = 0, grp_buf = 0, grp_cc = 0;
    int cur_nt = -1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.tiles_n;
      int mt = tile / p.tiles_n;
      const int twi = mt % p.tiles_w; mt /= p.tiles_w;
      const int thi = mt % p.tiles_h; mt /= p.tiles_h;
      const int tti = mt % p.tiles_t;
      const int tbi = mt / p.tiles_t;
      const int ho = thi * p.TH + th, wo = twi * p.TW + tw;
      const int t = tti * p.TT + tl, n = tbi * p.TB + nl;
      const bool valid = (nl < p.TB) && (ho < p.Ho) && (wo < p.Wo) && (t < p.To) && (n < p.N);
      const size_t pos = ((size_t)(n * p.To + t) * p.Ho + ho) * p.Wo + wo;
      size_t rpos = pos;
      if (res_mode == 2)
        rpos = ((size_t)(n * p.To + t) * (p.Ho >> 1) + (ho >> 1)) * (p.Wo >> 1) + (wo >> 1);
      const int nbase = nt * BN;
      const int ncols = min(BN, Cout - nbase);     // live output columns of this tile

      // scale / bias of this column tile -> smem (readers of the previous values are past their last
      // named barrier); skipped while consecutive tiles share the column tile
      if (nt != cur_nt) {
        cur_nt = nt;
        for (int j = ep_tid; j < BN; j += EPI_THREADS) {
          const int c = nbase + j;
          s_scale[j] = (p.scale && c < Cout) ? __ldg(p.scale + c) : 1.f;
          s_bias[j] = (p.bias && c < Cout) ? __ldg(p.bias + c) : 0.f;
        }
        epi_bar_sync();
      }

      mbar_wait(&tmem_full[as], aphase);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + as * BN + ((uint32_t)(lg * 32) << 16) + (uint32_t)(colw * part);
      uint32_t r[16];
      if (out_f32) tmem_ld_32x32b_x8_lo(taddr, r); else tmem_ld_32x32b_x16(taddr, r);
#pragma unroll 1
      for (int cc = 0; cc < ncols; cc += CW) {
        const int c0 = cc + colw * part;           // first column (inside the tile) this warp handles
        const int cbase = nbase + c0;
        const bool more = cc + CW < ncols;
        float v[16];
        if (out_f32) {
          // ---- fp32 output: this warp owns 8 columns = 32 bytes
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 s4 = lds_f4(sc_u32 + (uint32_t)(cc + 4 * q) * 4u), b4 = lds_f4(bi_u32 + (uint32_t)(cc + 4 * q) * 4u);
            v[4 * q] = fmaf(__uint_as_float(r[4 * q]), s4.x, b4.x);
            v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), s4.y, b4.y);
            v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), s4.z, b4.z);
            v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), s4.w, b4.w);
          }
          if (more) tmem_ld_32x32b_x8_lo(taddr + cc + CW, r);
          if (res_mode != 0 && valid) {
            const float* rp = reinterpret_cast<const float*>(p.residual) + rpos * p.res_ld + cbase;
            if (cbase + 8 <= Cout) {
#pragma unroll
              for (int j = 0; j < 8; j += 4) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rp + j));
                v[j] += q.x; v[j + 1] += q.y; v[j + 2] += q.z; v[j + 3] += q.w;
                if (split_out) {                                     // residual = hi + lo
                  const float4 ql = __ldg(reinterpret_cast<const float4*>(rp + p.res_lo_off + j));
                  v[j] += ql.x; v[j + 1] += ql.y; v[j + 2] += ql.z; v[j + 3] += ql.w;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (cbase + j < Cout) v[j] += __ldg(rp + j) + (split_out ? __ldg(rp + p.res_lo_off + j) : 0.f);
            }
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
          }
        } else {
          // ---- bf16 output: this warp owns 16 columns = 32 bytes
          // residual rows first: the global loads overlap the TMEM wait and the scale/bias reads
          uint4 resq[2];
          const bool res_on = res_mode != 0 && valid && !res_tma;
          bool res_vec = res_on && (cbase + 16 <= Cout);
          const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(p.residual) + rpos * p.res_ld + cbase;
          if (res_vec) {
            resq[0] = __ldg(reinterpret_cast<const uint4*>(rp));
            resq[1] = __ldg(reinterpret_cast<const uint4*>(rp + 8));
          }
          if (res_tma) {
            // the producer warp prefetched this chunk's residual rows (zero-filled outside the tensor) into
            // the swizzled ring: read this thread's 32 bytes, then hand the slot back
            mbar_wait(&r_full[rslot], rphase);
            const uint32_t src = rbuf_u32 + (uint32_t)rslot * Cfg::C_BYTES + row_smem;
            resq[0] = lds_u4(src + q0);
            resq[1] = lds_u4(src + q1);
            res_vec = true;
            __syncwarp();
            if (lane == 0) mbar_arrive(&r_empty[rslot]);
            if (++rslot == p.nrbuf) { rslot = 0; rphase ^= 1; }
          }
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 s4 = lds_f4(sc_u32 + (uint32_t)(cc + 4 * q) * 4u), b4 = lds_f4(bi_u32 + (uint32_t)(cc + 4 * q) * 4u);
            v[4 * q] = fmaf(__uint_as_float(r[4 * q]), s4.x, b4.x);
            v[4 * q + 1] = fmaf(__uint_as_float(r[4 * q + 1]), s4.y, b4.y);
            v[4 * q + 2] = fmaf(__uint_as_float(r[4 * q + 2]), s4.z, b4.z);
            v[4 * q + 3] = fmaf(__uint_as_float(r[4 * q + 3]), s4.w, b4.w);
          }
          if (more) tmem_ld_32x32b_x16(taddr + cc + CW, r);
          if (res_vec) {
#pragma unroll
            for (int g4 = 0; g4 < 2; ++g4) {
              const uint32_t w4[4] = {resq[g4].x, resq[g4].y, resq[g4].z, resq[g4].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                v[8 * g4 + 2 * e] += __uint_as_float(w4[e] << 16);
                v[8 * g4 + 2 * e + 1] += __uint_as_float(w4[e] & 0xffff0000u);
              }
            }
          } else if (res_on) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (cbase + j < Cout) v[j] += __bfloat162float(rp[j]);
          }
        }

        // staging buffers: the TMA stores that used this group's buffers last must have read them out
        if (gi == 0) {
          grp_buf = buf_idx; grp_cc = cc;
          if (store_warp) {
            if (elect_one()) {
              if (inflight <= 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              else if (inflight == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            }
            __syncwarp();
          }
          epi_bar_sync();
        }
        const uint32_t buf = cbuf_u32 + (uint32_t)buf_idx * Cfg::C_BYTES;
        const uint32_t dst = buf + row_smem;
        buf_idx += bstep;
        if (buf_idx >= ncbuf) buf_idx = 0;
        if (out_f32) {
          if (split_out) {
            float lo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const float hi = round_to_tf32(v[j]); lo[j] = round_to_tf32(v[j] - hi); v[j] = hi; }
            const uint32_t dst_lo = dst + Cfg::C_BYTES;               // the lo chunk uses the next staging buffer
            sts_f4(dst_lo + q0, lo[0], lo[1], lo[2], lo[3]);
            sts_f4(dst_lo + q1, lo[4], lo[5], lo[6], lo[7]);
          } else if (p.round_tf32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = round_to_tf32(v[j]);
          }
          sts_f4(dst + q0, v[0], v[1], v[2], v[3]);
          sts_f4(dst + q1, v[4], v[5], v[6], v[7]);
        } else {
          uint32_t h[8];
          if (relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = pack_bf16x2_relu(v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) h[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          }
          sts_b4(dst + q0, h[0], h[1], h[2], h[3]);
          sts_b4(dst + q1, h[4], h[5], h[6], h[7]);
        }
        // group complete (or last chunk of the tile): generic-proxy smem writes -> visible to the async
        // proxy, one barrier, then one thread stores every chunk of the group and commits them together
        if (gi == cgroup - 1 || !more) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          epi_bar_sync();
          if (store_warp && elect_one()) {
            const int cw0 = twi * p.TW, ch0 = thi * p.TH, ct0 = tti * p.TT, cn0 = tbi * p.TB;
            if (split_out) {
              tma_store_5d(&tmC, buf, nbase + cc, cw0, ch0, ct0, cn0);
              tma_store_5d(&tmC, buf + Cfg::C_BYTES, nbase + cc + p.out_lo_off, cw0, ch0, ct0, cn0);
            } else {
              for (int g = 0; g <= gi; ++g)
                tma_store_5d(&tmC, cbuf_u32 + (uint32_t)(grp_buf + g) * Cfg::C_BYTES, nbase + grp_cc + g * CW, cw0, ch0, ct0, cn0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          // keep groups aligned in the buffer ring when a tile ends with a short group
          buf_idx = grp_buf + cgroup * bstep;
          if (buf_idx >= ncbuf) buf_idx = 0;
          gi = 0;
        } else {
          ++gi;
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (store_warp && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int encode_map(CUtensorMap* m, bool f32, int rank, const void* base, const uint64_t* dims,
                      const uint64_t* strides_bytes /*rank-1*/, const uint32_t* box, const uint32_t* estr) {
  PFN_encodeTiled enc = get_encode();
  DT_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  cuuint64_t d[5], s[4]; cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank,
                   const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DT_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu ..., box %u %u %u)",
               (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
               (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
  return 0;
}

// M tile = TB images x TT frames x TH x TW output positions (<= 128 rows).  Small feature maps (14x14 RoI
// heads, 25x42 res5) would waste a quarter of every 128-row MMA with purely spatial tiles; stacking frames /
// images fills the rows.  Among the shapes within 4 % of the best useful-row fraction a purely spatial tile
// wins (the fullest, then the widest); otherwise the widest stacked tile (longest runs per TMA box).
struct TileShape { int th, tw, tt, tb; };
static TileShape pick_tile(int Ho, int Wo, int To, int N, int max_w, int max_h, bool stack_t) {
  TileShape best = {8, 16, 1, 1};
  double best_eff = -1;
  long best_rank = -1;
  for (int pass = 0; pass < 2; ++pass)
    for (int tw = 1; tw <= 128 && tw <= max_w && tw <= Wo; ++tw)
      for (int th = 1; th * tw <= 128 && th <= max_h && th <= Ho; ++th) {
        const int rem = 128 / (tw * th);
        for (int tt = 1; tt <= rem && tt <= To && (tt == 1 || stack_t); ++tt)
          for (int tb = 1; tb * tt <= rem && tb <= N; ++tb) {
            const double tiles = (double)cdiv(Wo, tw) * cdiv(Ho, th) * cdiv(To, tt) * cdiv(N, tb);
            const double eff = (double)Ho * Wo * To * N / (tiles * 128.0);
            if (pass == 0) { if (eff > best_eff) best_eff = eff; continue; }
            if (eff < best_eff - 0.04) continue;
            const bool plain = tt == 1 && tb == 1;
            const long e = (long)(eff * 1e6);
            const long rank = plain ? (1L << 40) + e * 1000L + tw : (long)tw * 10000000L + e;
            if (rank > best_rank) { best_rank = rank; best = {th, tw, tt, tb}; }
          }
      }
  return best;
}

// Output map: dims (Cout, Wo, Ho, To, N) of the NDHWC result, box = one staged chunk (128 B of channels x one M tile).
static int encode_out_map(CUtensorMap* m, void* y, int out_f32, int Cout, int Wo, int Ho, int To, int N, int out_ld,
                          const TileShape& ts, bool time_major = false) {
  const uint64_t oesz = out_f32 ? 4 : 2;
  uint64_t d[5] = {(uint64_t)Cout, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)To, (uint64_t)N};
  const uint64_t frame = (uint64_t)out_ld * oesz * Wo * Ho;
  uint64_t st[4] = {(uint64_t)out_ld * oesz, (uint64_t)out_ld * oesz * Wo, time_major ? frame * N : frame,
                    time_major ? frame : frame * To};
  uint32_t b[5] = {(uint32_t)(128 / oesz), (uint32_t)ts.tw, (uint32_t)ts.th, (uint32_t)ts.tt, (uint32_t)ts.tb};
  uint32_t e[5] = {1, 1, 1, 1, 1};
  return encode_map(m, out_f32 != 0, 5, y, d, st, b, e);
}

template <int BN, bool TF32>
static int launch_conv(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                       const ConvKernelParams& p, int grid, cudaStream_t stream) {
  using Cfg = ConvCfg<BN>;
  static bool attr = false;
  if (!attr) {
    DT_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN, TF32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       Cfg::BUDGET));
    attr = true;
  }
  ConvKernelParams q = p;
  Cfg::split(p.split_out ? 1 : p.kT * p.kH * p.kW * p.kchunks * (p.split_in ? 3 : 1), p.nrbuf > 0, &q.nstages, &q.ncbuf,
             &q.nrbuf);
  q.cgroup = p.split_out ? 1 : q.ncbuf;             // one named-barrier pair / commit group per ring of chunks
  if (q.cgroup > 2 && BN < 256) q.cgroup = 2;       // ... but no longer than a tile (BN <= 128: two bf16 chunks)
  const int smem = Cfg::smem_bytes(q.nstages, q.ncbuf, q.nrbuf);
  DT_CHECK_ARG(q.nstages >= 2 && smem <= Cfg::BUDGET, "conv: smem split failed (%d stages, %d B)", q.nstages, smem);
  conv_tc_kernel<BN, TF32><<<grid, CONV_THREADS, smem, stream>>>(tmA, tmB, tmC, tmR, q);
  DT_CHECK_LAUNCH();
  return 0;
}

}  // namespace dt

using namespace dt;

extern "C" int dt_conv3d(const dt_conv_desc* d, const void* x, const void* w, const float* scale, const float* bias,
                         const void* residual, void* y, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DT_CHECK_ARG(d != nullptr, "dt_conv3d: null descriptor");
  DT_CHECK_ARG(d->dtype == DT_DTYPE_BF16 || d->dtype == DT_DTYPE_TF32, "dt_conv3d: dtype %d not in {BF16, TF32}", d->dtype);
  const bool tf32 = d->dtype == DT_DTYPE_TF32;
  const int esz = tf32 ? 4 : 2;
  const int BK = tf32 ? 32 : 64;
  DT_CHECK_ARG(d->N >= 1 && d->Ti >= 1 && d->Hi >= 1 && d->Wi >= 1 && d->Cin >= 1 && d->Cout >= 1,
               "dt_conv3d: bad input shape N=%d T=%d H=%d W=%d Cin=%d Cout=%d", d->N, d->Ti, d->Hi, d->Wi, d->Cin, d->Cout);
  DT_CHECK_ARG(d->kT >= 1 && d->kH >= 1 && d->kW >= 1 && d->sT >= 1 && d->sH >= 1 && d->sW >= 1 && d->pT >= 0 &&
                   d->pH >= 0 && d->pW >= 0, "dt_conv3d: bad filter geometry");
  const int To = (d->Ti + 2 * d->pT - d->kT) / d->sT + 1;
  const int Ho = (d->Hi + 2 * d->pH - d->kH) / d->sH + 1;
  const int Wo = (d->Wi + 2 * d->pW - d->kW) / d->sW + 1;
  DT_CHECK_ARG(To >= 1 && Ho >= 1 && Wo >= 1, "dt_conv3d: empty output (%d,%d,%d)", To, Ho, Wo);
  const int in_ld = d->in_ld > 0 ? d->in_ld : d->Cin;
  const int w_ld = d->w_ld > 0 ? d->w_ld : d->Cin;
  const int out_ld = d->out_ld > 0 ? d->out_ld : d->Cout;
  const int res_ld = d->res_ld > 0 ? d->res_ld : d->Cout;
  DT_CHECK_ARG((in_ld * esz) % 16 == 0 && (w_ld * esz) % 16 == 0,
               "dt_conv3d: channel strides must be multiples of 16 bytes (in_ld=%d, w_ld=%d, %d B/elem)", in_ld, w_ld, esz);
  DT_CHECK_ARG(in_ld >= d->Cin && w_ld >= d->Cin && out_ld >= d->Cout, "dt_conv3d: leading dims smaller than channels");
  DT_CHECK_ARG(x && w && y, "dt_conv3d: null tensor pointer");
  DT_CHECK_ARG(((uintptr_t)x % 16) == 0 && ((uintptr_t)w % 16) == 0 && ((uintptr_t)y % 16) == 0,
               "dt_conv3d: tensors must be 16-byte aligned");
  DT_CHECK_ARG(d->res_mode >= 0 && d->res_mode <= 2 && (d->res_mode == 0 || residual), "dt_conv3d: bad residual mode/pointer");
  DT_CHECK_ARG(d->res_mode != 2 || (Ho % 2 == 0 && Wo % 2 == 0), "dt_conv3d: upsample-add needs even output size, got %dx%d", Ho, Wo);
  const int out_f32 = d->out_f32;
  const int oesz = out_f32 ? 4 : 2;
  // vector stores need 16-byte aligned rows
  DT_CHECK_ARG((out_ld * oesz) % 16 == 0 && (d->res_mode == 0 || (res_ld * oesz) % 16 == 0),
               "dt_conv3d: out_ld/res_ld rows must be 16-byte multiples");

  // stacking frames inside one TMA box needs unit temporal stride (pointwise convs fold strides into the map)
  const bool pointwise = d->kT == 1 && d->kH == 1 && d->kW == 1 && d->pT == 0 && d->pH == 0 && d->pW == 0;
  const TileShape ts = pick_tile(Ho, Wo, To, d->N, pointwise ? 256 : 256 / d->sW, pointwise ? 256 : 256 / d->sH,
                                 pointwise || d->sT == 1);
  const int TH = ts.th, TW = ts.tw;
  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.N = d->N; p.To = To; p.Ho = Ho; p.Wo = Wo; p.Cout = d->Cout;
  p.kT = d->kT; p.kH = d->kH; p.kW = d->kW; p.pT = d->pT; p.pH = d->pH; p.pW = d->pW;
  p.sT = d->sT; p.sH = d->sH; p.sW = d->sW;
  p.kchunks = cdiv(d->Cin, BK);
  p.TH = TH; p.TW = TW; p.TT = ts.tt; p.TB = ts.tb;
  p.tiles_h = cdiv(Ho, TH); p.tiles_w = cdiv(Wo, TW); p.tiles_t = cdiv(To, ts.tt); p.tiles_b = cdiv(d->N, ts.tb);
  p.a_bytes = (uint32_t)(TH * TW * ts.tt * ts.tb) * 128u;
  p.scale = scale; p.bias = bias; p.residual = residual; p.res_mode = d->res_mode; p.res_ld = res_ld;
  p.relu = d->relu; p.out = y; p.out_ld = out_ld; p.out_f32 = out_f32; p.round_tf32 = d->out_round_tf32;
  // 3xTF32 split operands / outputs
  p.split_in = (d->x3 & 1) ? 1 : 0;
  p.split_out = (d->x3 & 2) ? 1 : 0;
  DT_CHECK_ARG(!p.split_in || (tf32 && d->Cin % BK == 0), "dt_conv3d: x3 inputs need DT_DTYPE_TF32 and Cin %% %d == 0 (Cin=%d)", BK, d->Cin);
  DT_CHECK_ARG(!p.split_out || (out_f32 && d->Cout % 32 == 0), "dt_conv3d: x3 outputs need fp32 and Cout %% 32 == 0 (Cout=%d)", d->Cout);
  p.a_lo_off = d->in_lo_off > 0 ? d->in_lo_off : in_ld / 2;
  p.b_lo_off = w_ld / 2;
  p.out_lo_off = d->out_lo_off > 0 ? d->out_lo_off : out_ld / 2;
  p.res_lo_off = d->res_lo_off > 0 ? d->res_lo_off : res_ld / 2;
  DT_CHECK_ARG(!p.split_in || (p.a_lo_off + d->Cin <= in_ld && p.b_lo_off >= d->Cin), "dt_conv3d: x3 lo halves do not fit the rows");
  DT_CHECK_ARG(!p.split_out || p.out_lo_off + d->Cout <= out_ld, "dt_conv3d: x3 output lo half does not fit the row");

  int BN = d->Cout >= 256 ? 256 : (d->Cout > 64 ? 128 : (d->Cout > 32 ? 64 : 32));
  if (d->Cout > 128 && d->Cout < 256) BN = 128;
  // bf16 same-shape residual: its chunks are prefetched by TMA into a shared-memory ring (coalesced 128-byte
  // rows instead of one 32-byte global load per thread).  The ring needs room, so these layers use 128-wide
  // column tiles (they are the K-light 1x1 expansions of the bottlenecks: HBM-bound, not MMA-bound).
  const bool res_tma = d->res_mode == 1 && !out_f32 && !tf32 && ((uintptr_t)residual % 16) == 0;
  if (res_tma && BN > 128) BN = 128;
  p.nrbuf = res_tma ? 1 : 0;                         // ring depth is chosen with the smem split at launch
  p.tiles_n = cdiv(d->Cout, BN);
  const long long total = (long long)p.tiles_b * p.tiles_t * p.tiles_h * p.tiles_w * p.tiles_n;
  DT_CHECK_ARG(total < (1ll << 31), "dt_conv3d: too many tiles");
  p.total_tiles = (int)total;

  // ---- tensor maps -------------------------------------------------------------
  // A: dims (C, W, H, T, N) of the NDHWC input.  Pointwise strided convs (1x1x1, stride s, no
  // padding) fold the stride into the global strides so no element-stride traversal is needed;
  // other strided convs use TMA element strides (box covers s*TW input columns, every s-th kept).
  CUtensorMap tmA, tmB;
  uint64_t dims[5], strides[4]; uint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  const uint64_t sC = (uint64_t)in_ld * esz, sW = sC * d->Wi, sH = sW * d->Hi, sT = sH * d->Ti;
  const uint64_t cdim = p.split_in ? (uint64_t)in_ld : (uint64_t)d->Cin;
  if (pointwise) {
    dims[0] = cdim; dims[1] = Wo; dims[2] = Ho; dims[3] = To; dims[4] = d->N;
    strides[0] = sC * d->sW; strides[1] = sW * d->sH; strides[2] = sH * d->sT; strides[3] = sT;
    box[0] = BK; box[1] = TW; box[2] = TH; box[3] = ts.tt; box[4] = ts.tb;
    p.sT = p.sH = p.sW = 1;
  } else {
    dims[0] = cdim; dims[1] = d->Wi; dims[2] = d->Hi; dims[3] = d->Ti; dims[4] = d->N;
    strides[0] = sC; strides[1] = sW; strides[2] = sH; strides[3] = sT;
    box[0] = BK; box[1] = (uint32_t)TW * d->sW; box[2] = (uint32_t)TH * d->sH; box[3] = ts.tt; box[4] = ts.tb;
    estr[1] = d->sW; estr[2] = d->sH;
    DT_CHECK_ARG(box[1] <= 256 && box[2] <= 256, "dt_conv3d: strided tile too large for a TMA box");
  }
  if (encode_map(&tmA, tf32, 5, x, dims, strides, box, estr)) return 1;
  {
    const int taps = d->kT * d->kH * d->kW;
    uint64_t wd[3] = {p.split_in ? (uint64_t)w_ld : (uint64_t)d->Cin, (uint64_t)d->Cout, (uint64_t)taps};
    uint64_t ws[2] = {(uint64_t)w_ld * esz, (uint64_t)w_ld * esz * d->Cout};
    uint32_t wb[3] = {(uint32_t)BK, (uint32_t)BN, 1};
    uint32_t we[3] = {1, 1, 1};
    if (encode_map(&tmB, tf32, 3, w, wd, ws, wb, we)) return 1;
  }
  CUtensorMap tmC, tmR;
  if (encode_out_map(&tmC, y, out_f32, p.split_out ? out_ld : d->Cout, Wo, Ho, To, d->N, out_ld, ts, d->out_time_major != 0)) return 1;
  if (res_tma) {
    if (encode_out_map(&tmR, const_cast<void*>(residual), 0, d->Cout, Wo, Ho, To, d->N, res_ld, ts)) return 1;
  } else {
    tmR = tmC;
  }
  int dev = 0, sms = 148;
  DT_CHECK_CUDA(cudaGetDevice(&dev));
  DT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
#define DT_LAUNCH(BNv)                                                                      \
  return tf32 ? launch_conv<BNv, true>(tmA, tmB, tmC, tmR, p, grid, stream) : launch_conv<BNv, false>(tmA, tmB, tmC, tmR, p, grid, stream)
  switch (BN) {
    case 256: DT_LAUNCH(256);
    case 128: DT_LAUNCH(128);
    case 64: DT_LAUNCH(64);
    default: DT_LAUNCH(32);
  }
#undef DT_LAUNCH
}


// conv1 of the ResNet bodies: 7x7 stride 2 pad 3 on a 3-channel image (lib/modeling/ResNet3D.py:258-261,
// ResNet.py).  With Cin = 3 a per-tap k-block would waste 61/64 of every MMA, so the taps of one filter
// ROW are packed into K instead: the image blob is channel-padded to Cp (8 bf16 / 4 fp32 = 16 bytes per
// pixel) and carries physical zero borders (3 rows top/bottom, 4 pixels left/right, written by
// dt_prep_clip), so the 7-pixel window of output column wo is ONE contiguous 112-byte run starting at
// padded pixel 2*wo + 1.  The A tensor map is an overlapping strided view
//   dim0 = 8 pixels x Cp (128 B, the 8th pixel meets zero weights), dim1 = wo (stride 2 pixels = 32 B),
//   dim2 = rows of one parity plane, dim3 = plane, dim4 = frames
// (the blob's padded rows are de-interleaved by parity, so the stride-2 row walk of filter row kh is a
// unit-stride box in plane kh & 1 starting at row ho + (kh >> 1); TMA element strides halve its throughput)
// and the conv is 7 k-blocks (one per filter row) of K = 128 bytes: 3*7/ (8*8) = 33 % useful MACs instead
// of 4.7 %.  w [7][Cout][8*Cp] (kw-major, channel-minor).  y [F, Hp/2, Wp/2, out_ld].
extern "C" int dt_conv1_7x7s2(const void* x_padded, int F, int Hp, int Wp, int Cp, const void* w, int Cout,
                              const float* scale, const float* bias, int relu, int dtype, int out_f32,
                              int out_round_tf32, void* y, int out_ld, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DT_CHECK_ARG(dtype == DT_DTYPE_BF16 || dtype == DT_DTYPE_TF32, "dt_conv1_7x7s2: bad dtype %d", dtype);
  const bool tf32 = dtype == DT_DTYPE_TF32;
  const int esz = tf32 ? 4 : 2;
  DT_CHECK_ARG(Cp * esz == 16, "dt_conv1_7x7s2: the blob must carry 16 bytes per pixel (Cp=%d, %d B/elem)", Cp, esz);
  DT_CHECK_ARG(F >= 1 && Hp >= 2 && Wp >= 2 && Hp % 2 == 0 && Wp % 2 == 0 && Cout >= 1 && Cout <= 64,
               "dt_conv1_7x7s2: bad shape F=%d Hp=%d Wp=%d Cout=%d", F, Hp, Wp, Cout);
  DT_CHECK_ARG(x_padded && w && y, "dt_conv1_7x7s2: null pointer");
  const int oesz = out_f32 ? 4 : 2;
  if (out_ld <= 0) out_ld = Cout;
  DT_CHECK_ARG((out_ld * oesz) % 16 == 0 && out_ld >= Cout, "dt_conv1_7x7s2: bad out_ld %d", out_ld);
  const int Ho = Hp / 2, Wo = Wp / 2;
  const int BKe = 128 / esz;                       // elements per k-block
  const TileShape ts = pick_tile(Ho, Wo, 1, 1, 256, 128, false);      // spatial tiles only
  const int TH = ts.th, TW = ts.tw;
  ConvKernelParams p;
  memset(&p, 0, sizeof(p));
  p.N = F; p.To = 1; p.Ho = Ho; p.Wo = Wo; p.Cout = Cout;
  p.kT = 1; p.kH = 7; p.kW = 1; p.sT = 1; p.sH = 1; p.sW = 1; p.pT = 0; p.pH = 0; p.pW = 0;
  p.row_planes = 1;
  p.kchunks = 1;
  p.TH = TH; p.TW = TW; p.TT = 1; p.TB = 1; p.tiles_h = cdiv(Ho, TH); p.tiles_w = cdiv(Wo, TW); p.tiles_t = 1; p.tiles_b = F;
  p.tiles_n = 1;
  p.a_bytes = (uint32_t)TH * TW * 128u;
  p.scale = scale; p.bias = bias; p.relu = relu; p.out = y; p.out_ld = out_ld; p.out_f32 = out_f32;
  p.round_tf32 = out_round_tf32;
  const long long total = (long long)F * p.tiles_h * p.tiles_w;
  DT_CHECK_ARG(total < (1ll << 31), "dt_conv1_7x7s2: too many tiles");
  p.total_tiles = (int)total;
  // padded rows are stored de-interleaved (dt_prep_clip row_planes): filter row kh of output row ho reads
  // padded row 2*ho + kh = plane (kh & 1), plane row ho + (kh >> 1) -> unit-stride boxes, the plane is dim 3
  const uint64_t pix = 16, row = (uint64_t)(Wp + 8) * pix, plane = row * ((Hp + 6) / 2), frame = 2 * plane;
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[5] = {(uint64_t)BKe, (uint64_t)Wo, (uint64_t)((Hp + 6) / 2), 2, (uint64_t)F};
    uint64_t strides[4] = {2 * pix, row, plane, frame};
    uint32_t box[5] = {(uint32_t)BKe, (uint32_t)TW, (uint32_t)TH, 1, 1};
    uint32_t estr[5] = {1, 1, 1, 1, 1};
    if (encode_map(&tmA, tf32, 5, (const char*)x_padded + pix, dims, strides, box, estr)) return 1;
  }
  {
    uint64_t wd[3] = {(uint64_t)BKe, (uint64_t)Cout, 7};
    uint64_t ws[2] = {128, (uint64_t)128 * Cout};
    uint32_t wb[3] = {(uint32_t)BKe, 64, 1};
    uint32_t we[3] = {1, 1, 1};
    if (encode_map(&tmB, tf32, 3, w, wd, ws, wb, we)) return 1;
  }
  CUtensorMap tmC;
  if (encode_out_map(&tmC, y, out_f32, Cout, Wo, Ho, 1, F, out_ld, ts)) return 1;
  int dev = 0, sms = 148;
  DT_CHECK_CUDA(cudaGetDevice(&dev));
  DT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  return tf32 ? launch_conv<64, true>(tmA, tmB, tmC, tmC, p, grid, stream) : launch_conv<64, false>(tmA, tmB, tmC, tmC, p, grid, stream);
}
