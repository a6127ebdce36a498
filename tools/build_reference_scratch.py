"""Scratch build of the reference package for the "reference estimators run unchanged on 'cuda'" GPU tests.

    python tools/build_reference_scratch.py            # /root/reference -> baseline/_ref/bhmm (+ 3 Cython extensions)

`pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` fails on this image (the
reference's versioneer.py uses configparser.SafeConfigParser, removed in Python 3.12), so the package directory is
copied UNMODIFIED into the git-ignored baseline/_ref/ and its three hot-path Cython extensions (setup.py:58-69) are
built in place, exactly like tests/golden/build_reference_min.py does for the golden fixtures.  baseline/_ref is NOT
gpurun-ignored: it travels to the GPU box, where /root/reference does not exist.  TEST / BASELINE INFRASTRUCTURE: nothing
under bhmm_b200/ imports it.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, 'baseline', '_ref')


def have_build():
    d = os.path.join(DEST, 'bhmm', 'hidden', 'impl_c')
    return os.path.isdir(d) and any(f.startswith('hidden.') and f.endswith('.so') for f in os.listdir(d))


def build(src='/root/reference', force=False):
    if have_build() and not force:
        return DEST
    if not os.path.isdir(os.path.join(src, 'bhmm')):
        raise RuntimeError('reference checkout not found at %s' % src)
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(os.path.join(src, 'bhmm'), os.path.join(DEST, 'bhmm'))
    shutil.copy(os.path.join(ROOT, 'tests', 'golden', 'build_reference_min.py'), os.path.join(DEST, 'build_min.py'))
    r = subprocess.run([sys.executable, 'build_min.py'], cwd=DEST, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0 or not have_build():
        raise RuntimeError('building the reference extensions failed:\n' + r.stdout[-3000:])
    shutil.rmtree(os.path.join(DEST, 'build'), ignore_errors=True)
    return DEST


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
