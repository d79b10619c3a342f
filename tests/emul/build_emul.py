"""Build the CPU-emulation library (TEST INFRASTRUCTURE ONLY; see cuda_emul.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libppsci_b200_emul.so")
SRC = os.path.join(ROOT, "paddlescience_b200", "csrc")


def build(force: bool = False) -> str:
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(HERE, "cuda_emul.h"),
                                                              os.path.join(ROOT, "include", "ppsci_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O2", "-DPPSCI_EMUL", "-x", "c++", os.path.join(SRC, "engine.cu"),
           "-I" + HERE, "-I" + os.path.join(ROOT, "include"), "-I" + SRC, "-shared", "-fPIC", "-pthread", "-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
