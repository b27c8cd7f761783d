// Pairwise (tube) IoU and greedy NMS as a 64x64-tile bitmask + on-device scan.
//
// Reference semantics (bit-exact; this file is compiled with -fmad=false and
// uses explicit _rn intrinsics so nothing is contracted):
//   lib/utils/cython_bbox.pyx:16-56   IoU with the '+1' convention.  The C that
//       Cython emits promotes every "+ 1" and the area products to fp64 (the
//       literal is written 1.0); box_area / iw / ih are rounded to fp32 when
//       stored.  iou_pair_ref() reproduces that sequence operation by operation.
//   lib/utils/boxes.py:60-69          tubes: sequential fp32 sum over T, / T.
//   lib/utils/cython_nms.pyx:37-87    2-D NMS: fp32 throughout, suppress ovr >= thr,
//       survivors returned in ascending ORIGINAL index.
//   lib/nms/py_cpu_nms_tubes.py:17-53 tube NMS: mean-over-T IoU, a box survives
//       iff ovT <= thr (NaN => suppressed), survivors in score order.
//   lib/nms/nms_kernel.cu:35-150      design lineage (64x64 tiles of u64 masks);
//       unlike it, nothing here allocates, copies to the host or synchronises.
#include "common.cuh"
#include "box_math.cuh"
#include "../../include/dt_b200.h"
#include <math_constants.h>

namespace dt {

// iou_pair_ref (cython_bbox.pyx:34-56, one pair, one frame) lives in box_math.cuh (shared with targets.cu)
// Mean-over-T IoU exactly as boxes.py:60-69 evaluates it.
__device__ __forceinline__ float tube_iou_ref(const float* __restrict__ b,
                                              const float* __restrict__ q, int T) {
  float acc = iou_pair_ref(b, q);
  for (int t = 1; t < T; ++t) acc = __fadd_rn(acc, iou_pair_ref(b + 4 * t, q + 4 * t));
  return __fdiv_rn(acc, (float)T);
}

__global__ void bbox_overlaps_kernel(const float* __restrict__ boxes, int n, int ldb,
                                     const float* __restrict__ query, int k, int ldq,
                                     int T, float* __restrict__ out, int ldo) {
  // x -> query index (contiguous in the output row), y -> box index
  const int kk = blockIdx.x * blockDim.x + threadIdx.x;
  const int nn = blockIdx.y * blockDim.y + threadIdx.y;
  if (kk >= k || nn >= n) return;
  float b[DT_MAX_T * 4], q[DT_MAX_T * 4];
  for (int c = 0; c < 4 * T; ++c) { b[c] = boxes[(size_t)nn * ldb + c]; q[c] = query[(size_t)kk * ldq + c]; }
  out[(size_t)nn * ldo + kk] = tube_iou_ref(b, q, T);
}

// ---------------------------------------------------------------- NMS ------
// IoU as the two NMS implementations compute it (fp32; areas precomputed).
__device__ __forceinline__ float nms_iou_frame(const float* a, float aarea,
                                               const float* b, float barea) {
  const float xx1 = fmaxf(a[0], b[0]), yy1 = fmaxf(a[1], b[1]);
  const float xx2 = fminf(a[2], b[2]), yy2 = fminf(a[3], b[3]);
  const float w = fmaxf(0.f, __fadd_rn(__fsub_rn(xx2, xx1), 1.f));
  const float h = fmaxf(0.f, __fadd_rn(__fsub_rn(yy2, yy1), 1.f));
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
}

__device__ __forceinline__ uint32_t float_sort_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending uint == ascending float
}

// One CTA per problem: descending-score order with ties broken by descending
// original index (== reversing a stable ascending argsort).  Bitonic in smem.
__global__ void nms_sort_kernel(const float* __restrict__ dets, int nmax, int ld, int T,
                                const int* __restrict__ counts, int npow2,
                                int* __restrict__ order /*[B,nmax]*/) {
  extern __shared__ unsigned long long keys[];
  const int b = blockIdx.x;
  const int n = counts ? min(counts[b], nmax) : nmax;
  const float* d = dets + (size_t)b * nmax * ld;
  for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
    unsigned long long key = 0ull;   // pads sort to the end (descending)
    if (i < n) key = ((unsigned long long)float_sort_key(d[(size_t)i * ld + 4 * T]) << 32) |
                     (unsigned long long)(uint32_t)(i + 1);
    keys[i] = key;
  }
  __syncthreads();
  for (int k = 2; k <= npow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < npow2; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = keys[i], c = keys[ixj];
          const bool desc = ((i & k) == 0);
          if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    order[(size_t)b * nmax + i] = (int)(uint32_t)(keys[i] & 0xffffffffull) - 1;
}

// grid (colblk, rowblk, B), 64 threads.  mask[b][i][cb] bit j set <=> box at sorted
// position cb*64+j is suppressed by the box at sorted position i.
template <int T_CT>
__global__ void nms_mask_kernel(const float* __restrict__ dets, int nmax, int ld, int T_rt,
                                const int* __restrict__ counts, const int* __restrict__ order,
                                float thresh, int cmp_mode,
                                unsigned long long* __restrict__ mask, int nw) {
  const int T = T_CT > 0 ? T_CT : T_rt;
  const int cb = blockIdx.x, rb = blockIdx.y, b = blockIdx.z;
  if (cb < rb) return;
  const int n = counts ? min(counts[b], nmax) : nmax;
  if (rb * 64 >= n || cb * 64 >= n) return;
  const float* d = dets + (size_t)b * nmax * ld;
  const int* ord = order + (size_t)b * nmax;
  __shared__ float cbox[64][DT_MAX_T * 4];
  __shared__ float carea[64][DT_MAX_T];
  const int tid = threadIdx.x;
  const int cj = cb * 64 + tid;
  if (cj < n) {
    const float* p = d + (size_t)ord[cj] * ld;
    for (int t = 0; t < T; ++t) {
      const float x1 = p[4 * t], y1 = p[4 * t + 1], x2 = p[4 * t + 2], y2 = p[4 * t + 3];
      cbox[tid][4 * t] = x1; cbox[tid][4 * t + 1] = y1; cbox[tid][4 * t + 2] = x2; cbox[tid][4 * t + 3] = y2;
      carea[tid][t] = __fmul_rn(__fadd_rn(__fsub_rn(x2, x1), 1.f), __fadd_rn(__fsub_rn(y2, y1), 1.f));
    }
  }
  __syncthreads();
  const int ri = rb * 64 + tid;
  if (ri >= n) return;
  float rbox[DT_MAX_T * 4], rarea[DT_MAX_T];
  {
    const float* p = d + (size_t)ord[ri] * ld;
    for (int t = 0; t < T; ++t) {
      rbox[4 * t] = p[4 * t]; rbox[4 * t + 1] = p[4 * t + 1]; rbox[4 * t + 2] = p[4 * t + 2]; rbox[4 * t + 3] = p[4 * t + 3];
      rarea[t] = __fmul_rn(__fadd_rn(__fsub_rn(rbox[4 * t + 2], rbox[4 * t]), 1.f),
                           __fadd_rn(__fsub_rn(rbox[4 * t + 3], rbox[4 * t + 1]), 1.f));
    }
  }
  unsigned long long bits = 0ull;
  const int jend = min(64, n - cb * 64);
  const int jstart = (cb == rb) ? tid + 1 : 0;
  for (int j = jstart; j < jend; ++j) {
    bool sup;
    if (cmp_mode == DT_NMS_2D_GE) {             // cython_nms.pyx:83-85
      const float ovr = nms_iou_frame(rbox, rarea[0], cbox[j], carea[j][0]);
      sup = (ovr >= thresh);
    } else {                                    // py_cpu_nms_tubes.py:35-51
      float ov = 0.f;
      for (int t = 0; t < T; ++t)
        ov = __fadd_rn(ov, nms_iou_frame(rbox + 4 * t, rarea[t], cbox[j] + 4 * t, carea[j][t]));
      ov = __fdiv_rn(ov, (float)T);
      sup = !(ov <= thresh);
    }
    if (sup) bits |= (1ull << j);
  }
  mask[((size_t)b * nmax + ri) * nw + cb] = bits;
}

// One CTA per problem walks the mask rows in score order.
//   per 64-block: thread 0 resolves the diagonal word serially in registers,
//   then the whole CTA ORs the rows of the newly kept boxes into remv[].
__global__ void nms_scan_kernel(const unsigned long long* __restrict__ mask, int nmax, int nw,
                                const int* __restrict__ counts, const int* __restrict__ order,
                                int out_order, int max_keep,
                                int* __restrict__ keep, int* __restrict__ num_keep,
                                unsigned char* __restrict__ flags /*[B,nmax] scratch*/) {
  extern __shared__ unsigned long long sm[];
  unsigned long long* remv = sm;            // [nw]
  unsigned long long* kept = sm + nw;       // [nw]
  __shared__ unsigned long long diag[64];
  __shared__ unsigned long long s_kept;
  __shared__ int s_total, s_stop;
  const int b = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const int n = counts ? min(counts[b], nmax) : nmax;
  const unsigned long long* m = mask + (size_t)b * nmax * nw;
  const int* ord = order + (size_t)b * nmax;
  int* out = keep + (size_t)b * nmax;
  const int nblk = (n + 63) / 64;
  for (int w = tid; w < nw; w += nth) { remv[w] = 0ull; kept[w] = 0ull; }
  if (tid == 0) { s_total = 0; s_stop = 0; }
  __syncthreads();
  // keep[:max_keep] truncates the RETURNED list (generate_proposals.py:108-110):
  // in score order that is an early exit; in index order all survivors are needed first.
  const int limit = (max_keep > 0) ? max_keep : 0x7fffffff;
  const int stop_at = (out_order == DT_NMS_ORDER_SCORE) ? limit : 0x7fffffff;
  for (int blk = 0; blk < nblk; ++blk) {
    if (tid < 64) {
      const int i = blk * 64 + tid;
      diag[tid] = (i < n) ? m[(size_t)i * nw + blk] : 0ull;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long r = remv[blk], k = 0ull;
      int total = s_total;
      const int cnt = min(64, n - blk * 64);
      for (int j = 0; j < cnt; ++j) {
        if (!((r >> j) & 1ull)) {
          k |= (1ull << j);
          r |= diag[j];
          if (++total >= stop_at) { s_stop = 1; break; }
        }
      }
      s_kept = k; s_total = total; kept[blk] = k;
    }
    __syncthreads();
    if (s_stop) break;
    const unsigned long long k = s_kept;
    const int rem = nblk - blk - 1;
    for (int idx = tid; idx < 64 * rem; idx += nth) {
      const int j = idx / rem, w = blk + 1 + idx % rem;
      if ((k >> j) & 1ull) {
        const unsigned long long v = m[(size_t)(blk * 64 + j) * nw + w];
        if (v) atomicOr(&remv[w], v);
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int total = min(s_total, limit);
  if (tid == 0) num_keep[b] = total;
  if (out_order == DT_NMS_ORDER_SCORE) {
    // rank of sorted position p = popcount of kept bits below it
    for (int w = tid; w < nblk; w += nth) {
      int base = 0;
      for (int x = 0; x < w; ++x) base += __popcll(kept[x]);
      unsigned long long k = kept[w];
      while (k) {
        const int j = __ffsll((long long)k) - 1;
        k &= k - 1;
        out[base++] = ord[w * 64 + j];
      }
    }
  } else {
    // ascending original index: scatter flags, then an ordered block compaction
    unsigned char* f = flags + (size_t)b * nmax;
    for (int i = tid; i < n; i += nth) f[i] = 0;
    __syncthreads();
    for (int w = tid; w < nblk; w += nth) {
      unsigned long long k = kept[w];
      while (k) { const int j = __ffsll((long long)k) - 1; k &= k - 1; f[ord[w * 64 + j]] = 1; }
    }
    __syncthreads();
    __shared__ int s_warp[32];
    __shared__ int s_base;
    if (tid == 0) s_base = 0;
    __syncthreads();
    const int lane = tid & 31, wid = tid >> 5, nwarp = (nth + 31) >> 5;
    for (int start = 0; start < n; start += nth) {
      const int i = start + tid;
      const int v = (i < n) ? f[i] : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, v);
      const int pre = __popc(bal & ((1u << lane) - 1));
      if (lane == 0) s_warp[wid] = __popc(bal);
      __syncthreads();
      int off = s_base;
      for (int x = 0; x < wid; ++x) off += s_warp[x];
      if (v && off + pre < limit) out[off + pre] = i;
      __syncthreads();
      if (tid == 0) { int s = 0; for (int x = 0; x < nwarp; ++x) s += s_warp[x]; s_base += s; }
      __syncthreads();
    }
  }
}

static int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace dt

using namespace dt;

extern "C" int dt_bbox_overlaps(const float* boxes, int n, int ldb, const float* query, int k,
                                int ldq, int T, float* out, int ldo, void* stream) {
  DT_CHECK_ARG(T >= 1 && T <= DT_MAX_T, "dt_bbox_overlaps: T=%d outside [1,%d]", T, DT_MAX_T);
  DT_CHECK_ARG(n >= 0 && k >= 0 && ldb >= 4 * T && ldq >= 4 * T && ldo >= k,
               "dt_bbox_overlaps: bad shape n=%d k=%d ldb=%d ldq=%d ldo=%d T=%d", n, k, ldb, ldq, ldo, T);
  if (n == 0 || k == 0) return 0;
  DT_CHECK_ARG(boxes && query && out, "dt_bbox_overlaps: null pointer");
  dim3 blk(32, 8), grd(cdiv(k, 32), cdiv(n, 8));
  bbox_overlaps_kernel<<<grd, blk, 0, (cudaStream_t)stream>>>(boxes, n, ldb, query, k, ldq, T, out, ldo);
  DT_CHECK_LAUNCH();
  return 0;
}

extern "C" int dt_nms_workspace_bytes(int batch, int nmax, size_t* bytes) {
  DT_CHECK_ARG(batch >= 0 && nmax >= 0 && bytes, "dt_nms_workspace_bytes: bad args");
  const size_t nw = (size_t)cdiv(nmax, 64);
  size_t b = 0;
  b += align_up((size_t)batch * nmax * sizeof(int), 256);                       // order
  b += align_up((size_t)batch * nmax * nw * sizeof(unsigned long long), 256);   // mask
  b += align_up((size_t)batch * nmax, 256);                                     // flags
  *bytes = b;
  return 0;
}

extern "C" int dt_nms_batched(const float* dets, int batch, int nmax, int ld, int T,
                              const int* counts, float thresh, int cmp_mode, int out_order,
                              int max_keep, int* keep, int* num_keep, void* workspace,
                              size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  DT_CHECK_ARG(T >= 1 && T <= DT_MAX_T, "dt_nms_batched: T=%d outside [1,%d]", T, DT_MAX_T);
  DT_CHECK_ARG(batch >= 0 && nmax >= 0 && ld >= 4 * T + 1,
               "dt_nms_batched: bad shape batch=%d nmax=%d ld=%d T=%d (need ld >= 4T+1)", batch, nmax, ld, T);
  DT_CHECK_ARG(nmax <= DT_NMS_MAX_BOXES, "dt_nms_batched: nmax=%d exceeds %d", nmax, DT_NMS_MAX_BOXES);
  DT_CHECK_ARG(cmp_mode == DT_NMS_2D_GE || cmp_mode == DT_NMS_TUBE_GT, "dt_nms_batched: bad cmp_mode %d", cmp_mode);
  DT_CHECK_ARG(cmp_mode == DT_NMS_TUBE_GT || T == 1, "dt_nms_batched: the 2-D '>=' path takes T=1 (nms_wrapper.py:53-57)");
  DT_CHECK_ARG(out_order == DT_NMS_ORDER_SCORE || out_order == DT_NMS_ORDER_INDEX, "dt_nms_batched: bad out_order %d", out_order);
  if (batch == 0) return 0;
  DT_CHECK_ARG(keep && num_keep, "dt_nms_batched: null output");
  if (nmax == 0) { DT_CHECK_CUDA(cudaMemsetAsync(num_keep, 0, sizeof(int) * batch, stream)); return 0; }
  DT_CHECK_ARG(dets && workspace, "dt_nms_batched: null pointer");
  size_t need; dt_nms_workspace_bytes(batch, nmax, &need);
  DT_CHECK_ARG(workspace_bytes >= need, "dt_nms_batched: workspace %zu < %zu bytes", workspace_bytes, need);
  const int nw = cdiv(nmax, 64);
  char* ws = (char*)workspace;
  int* order = (int*)ws; ws += align_up((size_t)batch * nmax * sizeof(int), 256);
  unsigned long long* mask = (unsigned long long*)ws; ws += align_up((size_t)batch * nmax * nw * sizeof(unsigned long long), 256);
  unsigned char* flags = (unsigned char*)ws;

  const int npow2 = next_pow2(nmax);
  const size_t sort_smem = (size_t)npow2 * sizeof(unsigned long long);
  static DynSmemGrant grant;
  DT_CHECK_CUDA(grant_dyn_smem(nms_sort_kernel, (int)(next_pow2(DT_NMS_MAX_BOXES) * sizeof(unsigned long long)), &grant));
  nms_sort_kernel<<<batch, npow2 >= 1024 ? 1024 : (npow2 < 64 ? 64 : npow2), sort_smem, stream>>>(
      dets, nmax, ld, T, counts, npow2, order);
  DT_CHECK_LAUNCH();
  dim3 grd(nw, nw, batch);
  if (T == 1)
    nms_mask_kernel<1><<<grd, 64, 0, stream>>>(dets, nmax, ld, T, counts, order, thresh, cmp_mode, mask, nw);
  else if (T == 3)
    nms_mask_kernel<3><<<grd, 64, 0, stream>>>(dets, nmax, ld, T, counts, order, thresh, cmp_mode, mask, nw);
  else
    nms_mask_kernel<0><<<grd, 64, 0, stream>>>(dets, nmax, ld, T, counts, order, thresh, cmp_mode, mask, nw);
  DT_CHECK_LAUNCH();
  nms_scan_kernel<<<batch, 256, 2 * nw * sizeof(unsigned long long), stream>>>(
      mask, nmax, nw, counts, order, out_order, max_keep, keep, num_keep, flags);
  DT_CHECK_LAUNCH();
  return 0;
}
