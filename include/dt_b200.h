/* dt_b200.h — C ABI of libdt_b200.so, the B200 (sm_100a) hot path behind the
 * DetectAndTrack cfg / model-builder / tools surface.
 *
 * Conventions (all entry points):
 *   - plain C types only; every data pointer is a DEVICE pointer owned by the
 *     caller unless the parameter is documented as host; nothing is allocated,
 *     freed or synchronised inside; work is enqueued on `stream` (a
 *     cudaStream_t passed as void*, NULL = legacy default stream).
 *   - returns 0 on success, non-zero on error; dt_last_error() then holds the
 *     message (per host thread).  Python raises RuntimeError with that text,
 *     matching the reference's CAFFE_ENFORCE -> RuntimeError behaviour
 *     (/root/reference/tests/test_zero_even_op.py:41-44).
 *   - boxes are fp32 rows [x1,y1,x2,y2]*T (+score) exactly as the reference
 *     lays them out (lib/utils/boxes.py:26-57); integer outputs are int32.
 *
 * The reference has one C-ABI precedent for this path,
 *   void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num,
 *             int boxes_dim, float nms_overlap_thresh, int device_id);
 *   (/root/reference/lib/nms/gpu_nms.hpp:3-9): host pointers, internal
 *   malloc/memcpy/free.  dt_nms_batched replaces it with device pointers, a
 *   stream, a batch dimension and both reference comparators.
 * Everything else on the path is a Caffe2 Operator (C++ class ABI, un-vendored);
 * each function below cites the operator / python function it replaces.
 */
#ifndef DT_B200_H_
#define DT_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DT_B200_ABI_VERSION 1
#define DT_MAX_T 8              /* frames per tube supported by the box kernels */
#define DT_NMS_MAX_BOXES 8192   /* per problem */
#define DT_LSA_MAX_DIM 224      /* max(prev, cur) detections per frame pair */

/* NMS comparator / output order */
#define DT_NMS_2D_GE 0      /* lib/utils/cython_nms.pyx:83-85      suppress ovr >= thr (T must be 1) */
#define DT_NMS_TUBE_GT 1    /* lib/nms/py_cpu_nms_tubes.py:49-51   keep mean-IoU <= thr              */
#define DT_NMS_ORDER_SCORE 0  /* survivors in descending-score order (py_cpu_nms_tubes) */
#define DT_NMS_ORDER_INDEX 1  /* survivors in ascending input index (cython_nms np.where) */

const char* dt_last_error(void);
int dt_abi_version(void);

/* ---- boxes.cu ----------------------------------------------------------- */

/* lib/utils/boxes.py:60-69 bbox_overlaps -> lib/utils/cython_bbox.pyx:16-56.
 * boxes [n, ldb], query [k, ldq] (first 4*T columns used), out [n, ldo] =
 * mean over T of the '+1' IoU.  Bit-exact with the compiled reference. */
int dt_bbox_overlaps(const float* boxes, int n, int ldb, const float* query, int k, int ldq,
                     int T, float* out, int ldo, void* stream);

/* lib/core/nms_wrapper.py:49-70 nms / tube_nms, batched.
 * dets [batch, nmax, ld]: columns 0..4T-1 boxes, column 4T score.
 * counts [batch] (device) valid rows per problem, NULL = nmax.
 * keep [batch, nmax] receives indices into the problem's rows, num_keep [batch].
 * max_keep > 0 truncates the returned list (generate_proposals.py:108-110).
 * Ties in score: descending score, then descending index (the reference's
 * order is undefined under ties). */
int dt_nms_workspace_bytes(int batch, int nmax, size_t* bytes /*host out*/);
int dt_nms_batched(const float* dets, int batch, int nmax, int ld, int T, const int* counts,
                   float thresh, int cmp_mode, int out_order, int max_keep, int* keep,
                   int* num_keep, void* workspace, size_t workspace_bytes, void* stream);

/* ---- lsa.cu ------------------------------------------------------------- */

/* scipy.optimize.linear_sum_assignment as called at
 * lib/core/tracking_engine.py:237, batched: cost [batch, dmax, ldc] fp32
 * (rows = previous frame, cols = current frame), nrows/ncols [batch] device.
 * matches [batch, dmax]: matches[q] = assigned row p, or -1
 * (tracking_engine.py:229-246).  status [batch] (may be NULL): 1 = infeasible.
 * Indices equal scipy >= 1.6 bit-for-bit (see oracle/lsa.py). */
int dt_lsa_batched(const float* cost, int batch, int dmax, int ldc, const int* nrows,
                   const int* ncols, int* matches, int* status, void* stream);

/* lib/core/tracking_engine.py:158-246 fused: cost = weight * (1 - IoU) between
 * frame f-1 and frame f, then the assignment.  frames [nframes, dmax, ld]
 * (4*T box columns first), counts [nframes]; frame 0 and frames with
 * is_start[f] != 0 (may be NULL) get all -1 (first frame of a video, :283-285). */
int dt_match_frames(const float* frames, int nframes, int dmax, int ld, int T, const int* counts,
                    const unsigned char* is_start, float weight, int* matches, int* status,
                    void* stream);

/* lib/core/tracking_engine.py:272-350 id propagation.  video_first [nvideos]
 * (device) = index of the first frame of each video, ascending.  tracks
 * [nframes, dmax] (-1 beyond counts[f]).  ids: next_id++, and
 * `if next_id >= max_ids: next_id %= max_ids` (:339-345). */
int dt_assign_track_ids(const int* matches, const int* counts, const unsigned char* is_start,
                        int nframes, int dmax, const int* video_first, int nvideos, int first_id,
                        int max_ids, int* tracks, void* stream);

/* lib/core/tracking_engine.py:86-93,711-748: centre-frame selection (when
 * center_only), clip of the first box to [0,w]x[0,h], keep score >= conf and
 * (x2-x1)*(y2-y1) >= min_area.  boxes [nframes, dmax, ld] with score in
 * column 4*T; hw [nframes, 2] = (height, width); out [nframes, dmax, 4*T_out+1]
 * compacted in input order, counts_out [nframes], sel [nframes, dmax] (may be
 * NULL) = source row of each kept detection. */
int dt_prune_detections(const float* boxes, int nframes, int dmax, int ld, int T, int center_only,
                        const int* counts_in, const float* hw, float conf, float min_area,
                        float* out, int* counts_out, int* sel, void* stream);

/* ---- conv_tc.cu ---------------------------------------------------------- */

#define DT_DTYPE_BF16 0   /* bf16 activations/weights, fp32 accumulate (tcgen05 kind::f16)   */
#define DT_DTYPE_TF32 1   /* fp32 storage, tf32 multiply, fp32 accumulate (tcgen05 kind::tf32) */

/* Geometry + fused epilogue of one convolution (host struct, plain ints).
 * Replaces a Caffe2 Conv/ConvNd (engine=CUDNN) followed by AffineChannel[Nd]
 * (lib/ops/affine_channel_nd_op.cu:19-70), Sum and Relu
 * (lib/modeling/detector.py:410-436, lib/modeling/ResNet3D.py:21-101), the FPN
 * top-down UpsampleNearest+Sum (lib/modeling/FPN3D.py:186-222) and FC
 * (lib/modeling/head_builder.py:33-36; a 1x1x1 conv over W = #rows).
 *   x  [N, Ti, Hi, Wi, in_ld]   NDHWC (channels innermost), first Cin channels used
 *   w  [kT*kH*kW, Cout, w_ld]   tap-major, channels innermost (cross-correlation, as Caffe2)
 *   y  [N, To, Ho, Wo, out_ld]  first Cout channels written
 *   y = relu?( conv(x, w) * scale[c] + bias[c]  (+ residual) )
 * res_mode 0: none; 1: residual has y's shape (ld res_ld); 2: residual is
 * [N, To, Ho/2, Wo/2, res_ld] and is read at (ho/2, wo/2) (nearest 2x upsample).
 * Leading dims 0 => dense.  x/w rows must be 16-byte multiples, y/residual rows too. */
typedef struct dt_conv_desc {
  int N, Ti, Hi, Wi, Cin, Cout;
  int kT, kH, kW;
  int sT, sH, sW;
  int pT, pH, pW;
  int in_ld, w_ld, out_ld, res_ld;
  int dtype;      /* DT_DTYPE_* : type of x and w */
  int out_f32;    /* 1: y/residual fp32, 0: bf16 */
  int relu;
  int res_mode;
} dt_conv_desc;

int dt_conv3d(const dt_conv_desc* desc /*host*/, const void* x, const void* w, const float* scale,
              const float* bias, const void* residual, void* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DT_B200_H_ */
