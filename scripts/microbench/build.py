"""Builds the diagnostic tcgen05 probe kernels (scripts/microbench/microbench.cu) into their OWN shared library,
build/libdsb_microbench.so -- they are not part of the product library."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "danspeech_b200", "csrc")
OUT = os.path.join(ROOT, "build", "libdsb_microbench.so")


def build():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    srcs = [os.path.join(HERE, "microbench.cu"), os.path.join(HERE, "tma_bench.cu"), os.path.join(CSRC, "common.cu"),
            os.path.join(CSRC, "gemm_tc.cu")]      # gemm_tc.cu: make_tmap_bf16
    if os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", CSRC, "-o", OUT] + srcs
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build())
