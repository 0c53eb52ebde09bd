"""Build librodeo_cuda.so (sm_100a) in-tree with nvcc:  python -m rodeo_legacy_b200.build

One object per .cu (compiled in parallel), linked into rodeo_legacy_b200/lib/librodeo_cuda.so.
nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the repo
snapshot (it is git-ignored, not gpurun-ignored).
"""
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "librodeo_cuda.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=true", "-prec-div=true", "-prec-sqrt=true",      # IEEE div / sqrt, never fast-math
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
] + (["-DRODEO_DEBUG_RING"] if os.environ.get("RODEO_DEBUG_RING") else []) \
  + (["-DRODEO_DEBUG_POISON"] if os.environ.get("RODEO_DEBUG_POISON") else []) \
  + ([f"-DRODEO_BDW_MIN_BLOCKS={os.environ['RODEO_BDW_MIN_BLOCKS']}"] if os.environ.get("RODEO_BDW_MIN_BLOCKS") else []) \
  + ([f"-DRODEO_STAGED_WARPS={os.environ['RODEO_STAGED_WARPS']}"] if os.environ.get("RODEO_STAGED_WARPS") else []) \
  + ([f"-DRODEO_TPB_WARPS={os.environ['RODEO_TPB_WARPS']}"] if os.environ.get("RODEO_TPB_WARPS") else []) \
  + ([f"-DRODEO_TPB_MAXNREG={os.environ['RODEO_TPB_MAXNREG']}"] if os.environ.get("RODEO_TPB_MAXNREG") else []) \
  + ([f"-DRODEO_TPX_U={os.environ['RODEO_TPX_U']}"] if os.environ.get("RODEO_TPX_U") else []) \
  + ([f"-DRODEO_TPB_GVAR={os.environ['RODEO_TPB_GVAR']}"] if os.environ.get("RODEO_TPB_GVAR") else [])


def nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, ptxas_info=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.inl")) + [
        os.path.join(os.path.dirname(HERE), "include", "rodeo_cuda.h"), os.path.abspath(__file__)]
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    jobs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_info else []) + ["-c", src, "-o", obj]
            jobs.append((src, cmd))
    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources]

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, rc, out in ex.map(run, jobs):
                if verbose or rc or ptxas_info:
                    sys.stderr.write(f"--- {os.path.basename(src)}\n{out}\n")
                if rc:
                    raise RuntimeError(f"nvcc failed on {src}")
    if jobs or _stale(LIB, objs):
        cmd = [nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


def build_microbench():
    """tools/microbench: write_pattern (no-compute kernel writing the smoother's output pattern: what the
    memory system sustains for 288-byte runs of 65,536 interleaved streams), fp64_peak (sustained DFMA
    rate; MEASURED_PEAKS.json has no FP64 figure), the write_pattern2 / write_pattern3 store-scheme experiments and
    fp64_latency (dependent-issue latency of DFMA, 1/x, rsqrt).  bench.py runs the first two next to the kernels.
    Returns the binaries' paths."""
    root = os.path.dirname(HERE)
    outs = []
    for name in ("write_pattern", "fp64_peak", "write_pattern2", "write_pattern3", "fp64_latency"):
        src = os.path.join(root, "tools", "microbench", name + ".cu")
        out = src[:-3]
        if _stale(out, [src]):
            cmd = [nvcc(), "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-cudart", "static",
                   "-o", out, src]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"microbenchmark {name} failed to build")
        outs.append(out)
    return outs


if __name__ == "__main__":
    build_microbench()
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, ptxas_info="--ptxas" in sys.argv))
