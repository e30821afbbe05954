"""Build the product for sm_100a: `make -C dtv-utils_b200` (nvcc -gencode arch=compute_100a,code=sm_100a
-lineinfo; see Makefile).  Artefacts stay in-tree (lib/, bin/) so they travel to the GPU box."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path() -> str:
    return os.path.join(HERE, "lib", "libpapr_b200.so")


def cli_path() -> str:
    return os.path.join(HERE, "bin", "papr")


def _stale() -> bool:
    outs = [lib_path(), cli_path()]
    if not all(os.path.exists(o) for o in outs):
        return True
    newest_out = min(os.path.getmtime(o) for o in outs)
    srcs = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))]
    srcs += [os.path.join(HERE, "Makefile"), os.path.join(HERE, "..", "include", "papr_b200.h")]
    return any(os.path.getmtime(s) > newest_out for s in srcs)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA kernels + C ABI + CLI if sources are newer than the artefacts."""
    if force or _stale():
        if not _have_nvcc():
            if os.path.exists(lib_path()):
                return lib_path()  # GPU box without sources newer than the shipped build
            raise RuntimeError("nvcc not found and no prebuilt libpapr_b200.so")
        r = subprocess.run(["make", "-C", HERE, "-j4"], capture_output=True, text=True)
        if verbose or r.returncode != 0:
            print(r.stdout)
            print(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("build of libpapr_b200.so failed")
    return lib_path()


def _have_nvcc() -> bool:
    from shutil import which
    return which("nvcc") is not None or os.path.exists("/usr/local/cuda/bin/nvcc")
