#!/usr/bin/env python
"""Installs the UNMODIFIED reference (jdlaubrie/Kiltro, /root/reference/kiltro) into the git-ignored baseline/_ref/ so that
`bench.py --impl reference` and the `cpu_baseline` leg can time the reference's own CPU implementation of the hot path
on the GPU box's host cores (/root/reference does not exist there, baseline/_ref/ travels with the
repo snapshot).

The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref /root/reference` is not applicable;
this script is the install recipe instead:
  1. copy /root/reference/kiltro -> baseline/_ref/kiltro  (Python sources, byte for byte);
  2. rebuild its one native piece, FiniteElements/ComputeSparsityPattern.{pyx,h}, for this interpreter -- the shipped .so
     links libpython3.8 and cannot be imported here -- with the commands of the reference's own
     FiniteElements/Makefile:11-16 (cython --cplus, g++ -O3 -fPIC -shared).
Nothing under baseline/_ref/ is product code and nothing of it is tracked by git.

Usage: python baseline/make_ref.py [--force]
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
DEST = os.path.join(HERE, "_ref")


def installed():
    fe = os.path.join(DEST, "kiltro", "FiniteElements")
    return os.path.isfile(os.path.join(fe, "ComputeSparsityPattern.so")) and os.path.isfile(os.path.join(DEST, "kiltro", "FEMSolver.py"))


def install(force=False, verbose=True):
    """Returns "installed", "present" or "unavailable: <why>"."""
    if installed() and not force:
        return "present"
    if not os.path.isdir(os.path.join(REF, "kiltro")):
        return "unavailable: /root/reference is not present (the GPU box uses the prebuilt baseline/_ref)"
    import numpy as np
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    shutil.copytree(os.path.join(REF, "kiltro"), os.path.join(DEST, "kiltro"))
    fe = os.path.join(DEST, "kiltro", "FiniteElements")
    for f in ("ComputeSparsityPattern.so", "ComputeSparsityPattern.cpp"):       # shipped binaries: python3.8 ABI
        p = os.path.join(fe, f)
        if os.path.exists(p):
            os.remove(p)
    out = None if verbose else subprocess.DEVNULL
    try:
        subprocess.check_call([sys.executable, "-m", "cython", "--cplus", "-3", "ComputeSparsityPattern.pyx"], cwd=fe,
                              stdout=out, stderr=out)
        subprocess.check_call(
            ["g++", "-fPIC", "-shared", "-pthread", "-O3", "-fno-strict-aliasing", "-DNDEBUG", "-w",
             "ComputeSparsityPattern.cpp", "-o", "ComputeSparsityPattern.so", "-I.",
             "-I" + sysconfig.get_paths()["include"], "-I" + np.get_include()], cwd=fe, stdout=out, stderr=out)
    except Exception as e:                                                     # no cython / g++: say so, do not half-install
        shutil.rmtree(DEST, ignore_errors=True)
        return "unavailable: rebuilding the reference's Cython extension failed (%s)" % e
    with open(os.path.join(DEST, "INSTALL.txt"), "w") as f:
        f.write("copy of /root/reference/kiltro (unmodified Python sources) + ComputeSparsityPattern.so rebuilt from the "
                "reference's own .pyx/.h for %s by baseline/make_ref.py\n" % sys.version.split()[0])
    return "installed"


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))
