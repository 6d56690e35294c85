"""ORACLE build recipe (test infrastructure only).

  python oracle/build.py          -> oracle/_build/librank_oracle.so   (gcc on oracle/rank_oracle.c)
  python oracle/build.py --ref    -> additionally oracle/_ref/rank_cy*.so, the reference's own Cython module
                                     compiled from the source where it lies under /root/reference
                                     (torchreid/metrics/rank_cylib/rank_cy.pyx; recipe = the reference's
                                     root setup.py:21-27, default -O2, single thread).  Only possible in
                                     the build container; outputs are git-ignored and travel with gpurun.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_PYX = "/root/reference/torchreid/metrics/rank_cylib/rank_cy.pyx"


def build_c():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "librank_oracle.so")
    src = os.path.join(HERE, "rank_oracle.c")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", out, src])
    return out


def build_ref():
    """Cythonize + compile the reference .pyx into oracle/_ref/ (never into the repo history)."""
    if not os.path.exists(REF_PYX):
        return None
    out_dir = os.path.join(HERE, "_ref")
    os.makedirs(out_dir, exist_ok=True)
    existing = [f for f in os.listdir(out_dir) if f.startswith("rank_cy") and f.endswith(".so")]
    if existing:
        return os.path.join(out_dir, existing[0])
    import numpy as np
    work = os.path.join(out_dir, "_work")
    os.makedirs(work, exist_ok=True)
    c_file = os.path.join(work, "rank_cy.c")
    subprocess.check_call([sys.executable, "-m", "cython", "-3", REF_PYX, "-o", c_file])
    import sysconfig
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    out = os.path.join(out_dir, "rank_cy" + ext)
    inc = [sysconfig.get_paths()["include"], np.get_include()]
    cmd = ["gcc", "-O2", "-shared", "-fPIC", "-fwrapv", "-Wno-deprecated-declarations"]
    for i in inc:
        cmd += ["-I", i]
    cmd += ["-o", out, c_file]
    subprocess.check_call(cmd)
    shutil.rmtree(work, ignore_errors=True)
    return out


if __name__ == "__main__":
    print(build_c())
    if "--ref" in sys.argv:
        print(build_ref())
