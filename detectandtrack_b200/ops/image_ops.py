"""Device-tensor API over csrc/jpeg.cu: JPEG byte streams -> BGR uint8 frames in device memory (nvJPEG, dlopen'ed).
The reference decodes with cv2.imread on loader threads (lib/utils/image.py:51-63); see include/dt_b200.h."""
import ctypes as C

from .. import _lib as L


def jpeg_available():
    return bool(L.lib().dt_jpeg_available())


def jpeg_decode(streams, H, W, out=None, stream=None):
    """streams: list of n bytes objects (JPEG files of H x W pixels) -> uint8 cuda tensor [n, H, W, 3] in cv2.imread's BGR
    order (written into `out` if given).  GPU work goes to `stream` (a torch.cuda.Stream; default: the current one)."""
    torch = L.require_cuda()
    n = len(streams)
    if out is None:
        out = torch.empty((n, H, W, 3), dtype=torch.uint8, device='cuda')
    assert out.dtype == torch.uint8 and out.is_contiguous() and out.numel() == n * H * W * 3
    bufs = (C.c_char_p * n)(*[C.c_char_p(s) for s in streams])
    sizes = (C.c_size_t * n)(*[len(s) for s in streams])
    sp = C.c_void_p(stream.cuda_stream) if stream is not None else L.stream_ptr()
    L.call('dt_jpeg_decode', bufs, sizes, n, H, W, L.ptr(out), sp)
    return out
