"""In-tree build of libmvae_b200.so (sm_100a only, nvcc cross-compiles without a GPU).

    python -m mvae_b200.build [--force]

Objects are compiled in parallel (one nvcc per .cu) into mvae_b200/csrc/_build/ and linked with the static CUDA
runtime, so the library has no link-time dependency on libcuda / libcudart and can be dlopen()ed on a CPU-only box
(the driver entry point cuTensorMapEncodeTiled is resolved at run time through cudaGetDriverEntryPoint).
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libmvae_b200.so")
SOURCES = ["api.cu", "pm_kernels.cu", "pm_fwd.cu", "pm_bwd.cu", "manifold_ops.cu", "elbo_kernels.cu", "skinny_kernels.cu", "latent_fwd.cu", "latent_bwd.cu", "dp_step.cu", "iwae_kernels.cu", "input_kernels.cu", "conv_kernels.cu", "gemm_sm100.cu"]
HEADERS = ["mvae_common.cuh", "manifold_math.cuh", "pm_math.cuh", "pm_item.cuh", "latent_impl.cuh", "pm_params.cuh", "pm_kernels_impl.cuh", os.path.join("..", "..", "include", "mvae_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr"]


def _mtime(path):
    return os.path.getmtime(path) if os.path.exists(path) else 0.0


def _compile(src, force):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    deps = [os.path.join(CSRC, src)] + [os.path.join(CSRC, h) for h in HEADERS]
    if not force and _mtime(obj) > max(_mtime(d) for d in deps):
        return obj, False
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    with concurrent.futures.ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    changed = any(c for _, c in results)
    if changed or force or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print("linked", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
