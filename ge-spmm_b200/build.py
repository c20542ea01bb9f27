"""In-tree build of the product: ``make`` in this directory.

Outputs (git-ignored, shipped to the GPU box by gpurun):
  lib/libgespmm.so, spmm<EXT_SUFFIX>, bin/spmm_test
"""
import os
import subprocess
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "lib", "libgespmm.so")
CLI = os.path.join(HERE, "bin", "spmm_test")
EXT = os.path.join(HERE, "spmm" + sysconfig.get_config_var("EXT_SUFFIX"))


def artifacts():
    return [LIB, CLI, EXT]


def is_built():
    return all(os.path.exists(p) for p in artifacts())


def build(verbose=False, jobs=None):
    """Compile everything for sm_100a (nvcc cross-compiles without a GPU)."""
    jobs = jobs or max(1, os.cpu_count() or 2)
    cmd = ["make", "-C", HERE, "-j", str(jobs)]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building ge-spmm_b200 failed (see output above)")
    missing = [p for p in artifacts() if not os.path.exists(p)]
    if missing:
        raise RuntimeError("build finished but artifacts are missing: %s" % missing)
