"""TEST INFRASTRUCTURE: build the g++ CPU-emulation of the plain-CUDA kernels (libdktb200_emu.so)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "deep_kernel_transfer_b200", "csrc")
LIB = os.path.join(HERE, "libdktb200_emu.so")
SOURCES = ["conv_fp32.cu", "conv1_bwd_mma.cu", "bn_pool.cu", "head.cu", "gp.cu", "gp_large.cu", "gp_kernels.cu", "gp_spectral.cu", "optim.cu", "conv_generic.cu", "resnet_ops.cu", "episode_feed.cu"]


def build(force=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "dktb_common.cuh"), os.path.join(HERE, "cuda_emu.h"),
                   os.path.join(HERE, "emu_impl.cpp")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) < os.path.getmtime(LIB) for d in deps):
        return LIB
    objs, procs = [], []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for s in srcs + [os.path.join(HERE, "emu_impl.cpp")]:
        o = os.path.join(bdir, os.path.basename(s).rsplit(".", 1)[0] + ".o")
        cmd = ["g++", "-std=c++20", "-O2", "-fPIC", "-x", "c++", "-DDKTB_EMU", "-I", HERE, "-I", CSRC, "-c", s, "-o", o]
        procs.append(subprocess.Popen(cmd))
        objs.append(o)
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("g++ failed")
    subprocess.check_call(["g++", "-shared", "-o", LIB] + objs + ["-lpthread"])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
