"""Builds libekf_b200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "ekf_b200.cu")
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))) + \
       [os.path.join(HERE, "..", "include", "ekf_b200.h")]
OUT_DIR = os.path.join(HERE, "lib")
OUT = os.path.join(OUT_DIR, "libekf_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "--fmad=true"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


HOST_SRCS = [os.path.join(HERE, "host", f) for f in ("EKF.cpp", "config_yaml.cpp", "new_features.cpp", "trace_yaml.cpp", "image_generator.cpp")]
SAMPLE_SRC = os.path.join(HERE, "..", "samples", "ekf_main.cpp")
SAMPLE_OUT = os.path.join(OUT_DIR, "ekf_sample")
HOST_OUT = os.path.join(OUT_DIR, "libekf_host.so")


def build_host(force=False):
    """g++ build of the C++ host side (the drop-in EKF class of include/EKF.h) and the sample driver, linked against
    libekf_b200.so.  Output: openekfmonoslam_b200/lib/ekf_sample."""
    build()
    deps = HOST_SRCS + [SAMPLE_SRC, os.path.join(HERE, "..", "include", "EKF.h"), os.path.join(HERE, "..", "include", "ekfb_cv_compat.hpp"),
                        os.path.join(HERE, "..", "include", "ImageGenerator.h"), OUT]
    if not force and os.path.exists(SAMPLE_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(SAMPLE_OUT) for d in deps):
        return SAMPLE_OUT
    link = ["-L" + OUT_DIR, "-lekf_b200", "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-shared", "-fPIC", "-o", HOST_OUT] + HOST_SRCS + link + ["-lz"])
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-o", SAMPLE_OUT, SAMPLE_SRC, "-lekf_host"] + link)
    return SAMPLE_OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
