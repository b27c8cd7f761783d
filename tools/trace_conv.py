"""Per-tile timeline of one conv_tc launch (CTA 0): builds a -DDT_CONV_TRACE copy of the library, runs one
layer and prints clock64 stamps of the producer / MMA issuer / epilogue per tile (tuning aid, not product).
    python tools/trace_conv.py "res2 64>64" [n]"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_trace_lib():
    from detectandtrack_b200 import build as B
    out = os.path.join(B.LIBDIR, 'libdt_b200_trace.so')
    objs = []
    for f in sorted(os.listdir(B.CSRC)):
        if not f.endswith('.cu'):
            continue
        obj = os.path.join(B.OBJ, f[:-3] + ('.trace.o' if f == 'conv_tc.cu' else '.o'))
        if f == 'conv_tc.cu':
            subprocess.check_call([B.NVCC] + B.ARCH + B.COMMON + ['-DDT_CONV_TRACE', '-c', os.path.join(B.CSRC, f), '-o', obj])
        objs.append(obj)
    subprocess.check_call([B.NVCC] + B.ARCH + ['-shared', '-Xcompiler', '-fPIC', '-o', out] + objs + ['-lcudart_static', '-ldl', '-lpthread', '-lrt'])
    return out


def main():
    if sys.argv[1] == 'build':
        print(build_trace_lib())
        return
    import torch
    from detectandtrack_b200 import _lib as L
    L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), 'libdt_b200_trace.so')
    from detectandtrack_b200.ops import conv as cv
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import bench_conv
    only = sys.argv[1]
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for (name, T, H, W, Cin, Cout, k, s, p) in bench_conv.LAYERS:
        if only not in name:
            continue
        x = torch.randn((nb, T, H, W, Cin), device='cuda').bfloat16()
        w = cv.pack_weight(torch.randn((Cout, Cin) + k) * 0.02, cv.BF16)
        sc = torch.ones(Cout, device='cuda'); bi = torch.zeros(Cout, device='cuda')
        y = cv.conv3d(x, w, k, s, p, sc, bi, relu=True, out_f32=False, dtype=cv.BF16)
        res = torch.randn_like(y) if name.endswith('+res') else None
        zero = (C.c_longlong * (64 * 16))()
        L.lib().dt_conv_trace_write.argtypes = [C.c_void_p]
        for _ in range(2):
            L.lib().dt_conv_trace_write(zero)
            cv.conv3d(x, w, k, s, p, sc, bi, res, 1 if res is not None else 0, relu=True, out_f32=False, dtype=cv.BF16, out=y)
        torch.cuda.synchronize()
        buf = (C.c_longlong * (64 * 16))()
        L.lib().dt_conv_trace_read.argtypes = [C.c_void_p]
        assert L.lib().dt_conv_trace_read(buf) == 0
        t0 = buf[0]
        print(name, 'columns: P.begin P.end M.begin M.end E.tile E.acc_ready E.slot_free E.staged E.done  (cycles since first stamp; epilogue = warp 3)')
        for i in range(24):
            print(i, ' '.join('%7d' % (buf[i * 16 + j] - t0) for j in (0, 1, 2, 3, 4, 5, 6, 7, 9)) + '  | wait_empty %6d wait_full %6d tma_issue %6d' % (buf[i * 16 + 10], buf[i * 16 + 11], buf[i * 16 + 12]))
        break


if __name__ == '__main__':
    main()
