"""Builds libbpbreid_sm100.so in-tree (bpbreid_b200/csrc -> bpbreid_b200/libbpbreid_sm100.so) with nvcc for sm_100a.
nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with gpurun."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libbpbreid_sm100.so")
SOURCES = ["api.cu", "tc_gemm.cu", "nn_kernels.cu", "head.cu", "head_lowres.cu", "loss.cu", "distance.cu", "rank.cu", "eval_kernels.cu", "tc_wgrad.cu", "bwd_kernels.cu", "loss_bwd.cu", "head_bwd.cu", "rerank.cu", "input_pipeline.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bpbreid_sm100.h")]
    return max(os.path.getmtime(f) for f in files if os.path.isfile(f))


def build(force=False, verbose=False):
    obj_dir = os.path.join(HERE, "_obj")
    os.makedirs(obj_dir, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _deps_mtime():
        return OUT
    headers_m = max(os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh")))

    def compile_one(src):
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        sp = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(sp), headers_m):
            return obj, ""
        r = subprocess.run([NVCC, *FLAGS, "-c", sp, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    objs = [o for o, _ in results]
    r = subprocess.run([NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
