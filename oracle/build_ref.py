"""Build the reference's own Cython kernels into oracle/_ref/ (TEST INFRASTRUCTURE, not product).

Compiles the three reference modules *where they lie* under /root/reference/src_cpp
(hamiltonian_math.pyx, sparse_math.pyx, hilbert_math.pyx; module names as in the
reference's src_cpp/setup.py:36-38) into

    oracle/_ref/src/utils/{hamiltonian_math,sparse_math,hilbert_math}*.so

No reference source is copied into the repository: the generated C files go to a
temporary build directory that is deleted afterwards, only the shared objects stay
(oracle/_ref/ is git-ignored, but travels to the GPU box with the gpurun snapshot).

Environment drift handled here:
  * Cython 3: `prange(2**N)` in hilbert_math.pyx needs compiler directive cpow=True;
  * the default `gcc` wrapper may lack libgomp.spec -> force /usr/bin/gcc.

Usage:  python oracle/build_ref.py [--reference /root/reference] [--force]
"""
import argparse
import glob
import os
import shutil
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
MODULES = ("hamiltonian_math", "sparse_math", "hilbert_math")


def built():
    return all(glob.glob(os.path.join(OUT, "src", "utils", m + "*.so")) for m in MODULES)


def build(reference="/root/reference", force=False, quiet=True):
    if built() and not force:
        return True
    src_dir = os.path.join(reference, "src_cpp")
    if not os.path.isdir(src_dir):
        return False
    import numpy as np
    from Cython.Build import cythonize
    from setuptools import Extension, setup

    os.environ.setdefault("CC", "/usr/bin/gcc")
    os.environ.setdefault("LDSHARED", "/usr/bin/gcc -shared")
    os.makedirs(os.path.join(OUT, "src", "utils"), exist_ok=True)
    for d in (os.path.join(OUT, "src"), os.path.join(OUT, "src", "utils")):
        open(os.path.join(d, "__init__.py"), "a").close()
    tmp = tempfile.mkdtemp(prefix="naqs_ref_build_")
    cwd = os.getcwd()
    try:
        os.chdir(src_dir)  # cythonize wants relative sources; outputs go to build_dir
        exts = [Extension("src.utils." + m, [m + ".pyx"],
                          extra_compile_args=["-fopenmp", "-O2", "-w"],
                          extra_link_args=["-fopenmp"],
                          include_dirs=[np.get_include()]) for m in MODULES]
        exts = cythonize(exts, build_dir=os.path.join(tmp, "c"), quiet=quiet,
                         compiler_directives={"cpow": True, "language_level": "3str"})
        argv = ["build_ext", "--build-lib", OUT, "--build-temp", os.path.join(tmp, "o")]
        if quiet:
            argv.insert(0, "-q")
        setup(name="naqs_ref_kernels", ext_modules=exts, script_args=argv)
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)
    return built()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    ok = build(a.reference, a.force, quiet=False)
    print("oracle/_ref built:", ok, sorted(glob.glob(os.path.join(OUT, "src", "utils", "*.so"))))
    sys.exit(0 if ok else 1)
