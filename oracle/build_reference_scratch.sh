#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds a runnable copy of the *reference package*
# (mrrt.nufft) in a scratch directory OUTSIDE the repo so that
# tests/golden/make_golden.py can import it and freeze golden vectors.
# Nothing from /root/reference is copied into the repo.
#
#   usage: oracle/build_reference_scratch.sh [/tmp/refbuild]
#
# Steps: writable copy of the tree; a minimal stand-in for the
# un-vendored, unpinned third-party dependency `mrrt.utils` (oracle/mrrt_utils_shim.py);
# Cython extension built with the reference's flags (-ffast-math -fopenmp,
# setup.py:48-50,110) by a 10-line setup script because the reference's setup.py needs
# distutils; and the one-line SciPy fix `obj.p.H` -> `obj.p.conj().T` (_nufft.py:1505;
# `.H` was removed from SciPy sparse matrices).
set -e
DST=${1:-/tmp/refbuild}
HERE=$(cd "$(dirname "$0")" && pwd)
rm -rf "$DST" && mkdir -p "$DST"
cp -r /root/reference/mrrt "$DST"/
mkdir -p "$DST"/mrrt/utils
cp "$HERE"/mrrt_utils_shim.py "$DST"/mrrt/utils/__init__.py
cat > "$DST"/setup_min.py <<'PY'
import numpy
from setuptools import setup, Extension
from Cython.Build import cythonize
ext = Extension("mrrt.nufft._extensions._nufft_table",
    sources=["mrrt/nufft/_extensions/c/nufft_table.c", "mrrt/nufft/_extensions/_nufft_table.pyx"],
    include_dirs=["mrrt/nufft/_extensions/c", numpy.get_include()],
    extra_compile_args=["-ffast-math", "-fopenmp"], extra_link_args=["-fopenmp"])
setup(name="refbuild", ext_modules=cythonize([ext], language_level=2))
PY
cd "$DST"
CC=/usr/bin/gcc LDSHARED="/usr/bin/gcc -shared" python setup_min.py build_ext --inplace > build.log 2>&1
sed -i 's/obj\.p\.H \* xk/obj.p.conj().T * xk/' mrrt/nufft/_nufft.py
echo "reference scratch build ready: PYTHONPATH=$DST"
