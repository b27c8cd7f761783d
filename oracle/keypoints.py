"""Oracle: keypoint heat-map decoding (TEST INFRASTRUCTURE ONLY).

Restates lib/utils/keypoints.py:94-149 (heatmaps_to_keypoints), :210-218
(scores_to_probs) and lib/core/test.py:865-894 (keypoint_results' per-frame loop),
calling the installed OpenCV's INTER_CUBIC resize exactly as the reference does
(OpenCV is an un-vendored dependency: pinned 3.4.1, here 4.13 — parity unpinned).
``bilinear_upsample2x`` / ``deconv_k4s2p1`` restate the fixed ConvTranspose of
lib/modeling/detector.py:348-380 and the kps_score_lowres ConvTranspose
(lib/modeling/model_builder.py:848-856) with torch CPU fp32 (Caffe2 stand-in)."""
import numpy as np

F32 = np.float32


def scores_to_probs(scores):
    for c in range(scores.shape[0]):
        t = scores[c]
        m = t.max()
        scores[c] = np.exp(t - m) / np.sum(np.exp(t - m))
    return scores


def heatmaps_to_keypoints(maps, rois, num_keypoints=17, min_size=0):
    """maps (R, K, M, M) fp32, rois (R, 4) -> (R, 4, K) fp32: x, y, logit, prob."""
    import cv2
    offset_x, offset_y = rois[:, 0], rois[:, 1]
    widths = np.maximum(rois[:, 2] - rois[:, 0], 1)
    heights = np.maximum(rois[:, 3] - rois[:, 1], 1)
    widths_ceil, heights_ceil = np.ceil(widths), np.ceil(heights)
    maps = np.transpose(maps, [0, 2, 3, 1])
    xy = np.zeros((len(rois), 4, num_keypoints), dtype=F32)
    for i in range(len(rois)):
        if min_size > 0:
            rw = int(np.maximum(widths_ceil[i], min_size)); rh = int(np.maximum(heights_ceil[i], min_size))
        else:
            rw, rh = widths_ceil[i], heights_ceil[i]
        wc = widths[i] / rw
        hc = heights[i] / rh
        roi_map = cv2.resize(np.ascontiguousarray(maps[i]), (int(rw), int(rh)), interpolation=cv2.INTER_CUBIC)
        if roi_map.ndim == 2:
            roi_map = roi_map[:, :, None]
        roi_map = np.transpose(roi_map, [2, 0, 1])
        probs = scores_to_probs(roi_map.copy())
        w = roi_map.shape[2]
        for k in range(num_keypoints):
            pos = roi_map[k].argmax()
            x_int = pos % w
            y_int = (pos - x_int) // w
            xy[i, 0, k] = (x_int + 0.5) * wc + offset_x[i]
            xy[i, 1, k] = (y_int + 0.5) * hc + offset_y[i]
            xy[i, 2, k] = roi_map[k, y_int, x_int]
            xy[i, 3, k] = probs[k, y_int, x_int]
    return xy


def keypoint_results(heatmaps, ref_boxes, num_keypoints=17):
    """test.py:865-894: per frame t, decode channels [t*K, (t+1)*K) against box t; concat on last axis."""
    T = heatmaps.shape[1] // num_keypoints
    parts = [heatmaps_to_keypoints(heatmaps[:, t * num_keypoints:(t + 1) * num_keypoints],
                                   ref_boxes[:, 4 * t:4 * t + 4], num_keypoints) for t in range(T)]
    return np.concatenate(parts, axis=-1)


def upsample_filt(size):
    factor = (size + 1) // 2
    center = factor - 1 if size % 2 == 1 else factor - 0.5
    og = np.ogrid[:size, :size]
    return (1 - abs(og[0] - center) / factor) * (1 - abs(og[1] - center) / factor)


def bilinear_upsample2x(x):
    """detector.py:348-380 with up_scale=2: x (N, K, H, W) torch fp32 -> (N, K, 2H, 2W)."""
    import torch
    import torch.nn.functional as Fn
    K = x.shape[1]
    w = torch.zeros((K, K, 4, 4), dtype=torch.float32, device=x.device)
    f = torch.from_numpy(upsample_filt(4).astype(np.float32)).to(x.device)
    for k in range(K):
        w[k, k] = f
    return Fn.conv_transpose2d(x, w, None, stride=2, padding=1)


def deconv_k4s2p1(x, w, b):
    """Caffe2 ConvTranspose kernel 4 stride 2 pad 1; w is (Cin, Cout, 4, 4) like Caffe2/torch."""
    import torch.nn.functional as Fn
    return Fn.conv_transpose2d(x, w, b, stride=2, padding=1)


# ------------------------------------------------------------------ pose-PCK tracking cost
# lib/utils/keypoints.py:266-291 and lib/core/tracking_engine.py:113-129 ('pose-pck' entry of
# TRACKING.DISTANCE_METRICS, weight 0 in every shipped yaml).  Restated for the device kernel of the next round;
# pinned by tests/golden/pose_pck.npz (generated from the reference's own functions).
def compute_head_size(kps, kpt_names):
    """:266-274.  |head_top - head_bottom| + 1 in the dtype of kps (float32 poses stay float32)."""
    ht = kps[:2, kpt_names.index('head_top')]
    hb = kps[:2, kpt_names.index('head_bottom')]
    return np.linalg.norm(ht - hb) + 1


def pck_distance(kps_a, kps_b, kpt_names, dist_thresh=0.5):
    """:277-291.  1 - (fraction of joints closer than dist_thresh head sizes); kps_a supplies the head size."""
    head = compute_head_size(kps_a, kpt_names)
    normed = np.linalg.norm(kps_a[:2] - kps_b[:2], axis=0) / head
    match = normed < dist_thresh
    return 1.0 - np.sum(match) / match.size


def pairwise_kpt_distance(a, b, kpt_names):
    """tracking_engine.py:113-129: res[i, j] = pck_distance(a[i], b[j]) as float64 [len(a), len(b)]."""
    res = np.zeros((len(a), len(b)))
    for i in range(len(a)):
        for j in range(len(b)):
            res[i, j] = pck_distance(a[i], b[j], kpt_names)
    return res
