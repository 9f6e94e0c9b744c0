"""Build libpfh_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pfhubbenchmarks_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpfh_b200.so")
SOURCES = ["pfh_stencil.cu", "pfh_stencil3_tma.cu", "pfh_nl.cu", "pfh_scalar3_tma.cu", "pfh_sample.cu", "pfh_vec.cu", "pfh_solver.cpp", "pfh_comm.cpp", "pfh_capi.cpp", "pfh_tables.cpp"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    return os.path.join(cuda_home, "bin", "nvcc")


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "pfh_b200.h"))
    return out


def up_to_date():
    return os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in _deps())


def build(force=False, verbose=False):
    if not force and up_to_date():
        return LIB
    if force:
        os.environ["PFH_FORCE_BUILD"] = "1"
    try:
        return _compile_all(os.path.join(HERE, "build"), LIB, [], verbose)
    finally:
        os.environ.pop("PFH_FORCE_BUILD", None)


def build_user(source, tag, force=False):
    """Compile a copy of the library for one generated model (codegen.UserModel): _jit/pfh_user_<tag>.cuh is written and
    compiled in through -DPFH_USER_PROBLEM_HEADER; the result is _jit/libpfh_b200_<tag>.so (cached by content hash and by
    the mtimes of the library sources).  -Bsymbolic keeps its internal calls inside this copy when the stock library is
    loaded in the same process."""
    jit = os.path.join(HERE, "_jit")
    os.makedirs(jit, exist_ok=True)
    hdr = os.path.join(jit, "pfh_user_%s.cuh" % tag)
    lib = os.path.join(jit, "libpfh_b200_%s.so" % tag)
    if not os.path.exists(hdr) or open(hdr).read() != source:
        with open(hdr, "w") as f:
            f.write(source)
    if not force and os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in _deps() + [hdr]):
        return lib
    if force:
        os.environ["PFH_FORCE_BUILD"] = "1"
    try:
        return _compile_all(os.path.join(jit, "build_" + tag), lib,
                            ['-DPFH_USER_PROBLEM_HEADER="%s"' % hdr, "-DPFH_ONLY_USER"], False, ["-Xlinker", "-Bsymbolic"])
    finally:
        os.environ.pop("PFH_FORCE_BUILD", None)


def _compile_all(objdir, lib, defines, verbose, link_flags=()):
    """Compile and link under an exclusive file lock (every rank of a torchrun job builds the same model at the same
    time): the first process builds into a temporary name and renames it into place, the others wait and find it done."""
    import fcntl
    os.makedirs(objdir, exist_ok=True)
    with open(lib + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            deps = _deps() + [d.split('"')[1] for d in defines if d.startswith("-DPFH_USER_PROBLEM_HEADER=")]
            if os.path.exists(lib) and all(os.path.getmtime(lib) >= os.path.getmtime(d) for d in deps) and not os.environ.get("PFH_FORCE_BUILD"):
                return lib                           # another process built it while this one waited for the lock
            tmp = lib + ".tmp%d" % os.getpid()
            _compile_locked(objdir, tmp, defines, verbose, link_flags)
            os.replace(tmp, lib)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return lib


def _compile_locked(objdir, lib, defines, verbose, link_flags=()):
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + ARCH + list(defines)
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src):
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        cmd = [_nvcc()] + flags + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [_nvcc()] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl"] + list(link_flags)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
