"""Build oracle/_ref/: the reference's OWN Cython IoU / NMS compiled from where
the sources lie under /root/reference (TEST INFRASTRUCTURE ONLY).

  lib/utils/cython_bbox.pyx  -> oracle/_ref/cython_bbox*.so   (unmodified)
  lib/utils/cython_nms.pyx   -> oracle/_ref/cython_nms*.so    (np.int_t / np.int
        spelled np.intp_t / np.intp on a scratch copy in a temp dir: numpy 2 /
        Cython 3 removed the alias; the arithmetic is untouched)

No reference source is copied into the repo: only the built .so files land in
oracle/_ref/ (git-ignored, NOT gpurun-ignored, so they travel to the GPU box).
Runs only where /root/reference exists (this container); silently skipped
elsewhere.  Usage: python oracle/build_ref.py
"""
import os, re, shutil, subprocess, sys, sysconfig, tempfile

REF = '/root/reference/lib/utils'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '_ref')


def _compile(pyx_path, name, tmp):
    import numpy
    c_path = os.path.join(tmp, name + '.c')
    subprocess.check_call([sys.executable, '-m', 'cython', '-3', '--directive',
                           'language_level=2', pyx_path, '-o', c_path])
    so = os.path.join(OUT, name + sysconfig.get_config_var('EXT_SUFFIX'))
    subprocess.check_call(['gcc', '-O2', '-fPIC', '-shared', '-fno-fast-math',
                           '-ffp-contract=off', '-Wno-cpp', '-Wno-unused-function',
                           '-I' + sysconfig.get_paths()['include'],
                           '-I' + numpy.get_include(), c_path, '-o', so])
    return so


def build(force=False):
    if not os.path.isdir(REF):
        return False
    os.makedirs(OUT, exist_ok=True)
    open(os.path.join(OUT, '__init__.py'), 'a').close()
    ext = sysconfig.get_config_var('EXT_SUFFIX')
    if not force and all(os.path.exists(os.path.join(OUT, n + ext))
                         for n in ('cython_bbox', 'cython_nms')):
        return True
    with tempfile.TemporaryDirectory() as tmp:
        _compile(os.path.join(REF, 'cython_bbox.pyx'), 'cython_bbox', tmp)
        src = open(os.path.join(REF, 'cython_nms.pyx')).read()
        src = src.replace('np.int_t', 'np.intp_t').replace('dtype=np.int)', 'dtype=np.intp)')
        patched = os.path.join(tmp, 'cython_nms.pyx')
        open(patched, 'w').write(src)
        _compile(patched, 'cython_nms', tmp)
    return True


if __name__ == '__main__':
    print('built' if build(force='--force' in sys.argv) else 'reference not present; skipped')
