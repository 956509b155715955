"""Builds eth-lc-plonky2_b200/libplonky2_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO = os.environ.get("ENG_SO", os.path.join(HERE, "libplonky2_b200.so"))  # ENG_SO: development A/B builds
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["engine.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--use_fast_math", "-Xptxas", "-v", "-ccbin", "/usr/bin/g++",
]


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "plonky2_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    # torchrun starts one process per GPU: exactly one of them compiles, the others wait for the lock and find the library
    # fresh.  The freshness check happens UNDER the lock, and nvcc writes to a temporary path that is renamed onto SO, so
    # a rank arriving mid-build can neither skip the lock on a fresh-looking mtime nor dlopen a half-written file.
    import fcntl
    with open(SO + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not needs_build():
            return SO
        return _build_locked(verbose)


def _build_locked(verbose):
    tmp = "%s.tmp.%d" % (SO, os.getpid())
    cmd = [NVCC] + FLAGS + os.environ.get("ENG_NVCC_EXTRA", "").split() + ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode == 0:
        os.replace(tmp, SO)          # atomic swap
    elif os.path.exists(tmp):
        os.remove(tmp)
    log = os.path.join(HERE, "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout[-4000:])
    return SO


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
