"""Build the CUDA library in-tree: vicres_b200/libvicres.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libvicres.so")
SOURCES = ["vicres_capi.cu", "vicres_route.cu", "vicres_atmos.cu", "vicres_files.cpp"]
DEPS = ["vicres_capi.cu", "vicres_route.cu", "vicres_files.cpp", "vicres_atmos.cu", "vic_atmos.cuh", "vic_cr_trig.cuh", "../../include/vicres_atmos.h", "../../include/vicres_files.h", "vic_physics.cuh", "vic_aggregate.cuh", "vic_host_prep.h",
        "../../include/vicres.h", "../../include/vicres_route.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-DVR_PHASE_BARRIERS"]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, fmad=False, extra=(), verbose=False):
    """Compile libvicres.so with nvcc.  -fmad=false keeps device arithmetic in the reference's
    operation order (no FMA contraction), see DESIGN.md section 5."""
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-fmad=%s" % ("true" if fmad else "false")] + list(extra)
    cmd += ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose=True)
    print(LIB)
