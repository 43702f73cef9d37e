"""TEST INFRASTRUCTURE — builds oracle/_build/libsched_emul.so: the product's host-only schedule files
(magma_b200/csrc/vit_train.cu and gptj_sched.cu, compiled as plain C++ — they contain no kernels and no CUDA calls, see
magma_b200/csrc/sched_rt.h) linked against oracle/cabi_emul.cpp, the CPU emulation of the primitive C-ABI operators.
tests/test_sched_emul_cpu.py loads it to dry-run the schedules against the oracle. Nothing in magma_b200/ uses it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT_DIR = os.path.join(ROOT, "oracle", "_build")
OUT = os.path.join(OUT_DIR, "libsched_emul.so")
SCHEDULES = [os.path.join(ROOT, "magma_b200", "csrc", "vit_train.cu"),
             os.path.join(ROOT, "magma_b200", "csrc", "gptj_sched.cu")]
EMUL = os.path.join(ROOT, "oracle", "cabi_emul.cpp")
EMUL_MODELS = os.path.join(ROOT, "oracle", "cabi_emul_models.cpp")  # model-level entries, delegating to the schedules
DEPS = SCHEDULES + [EMUL, EMUL_MODELS, os.path.join(ROOT, "magma_b200", "csrc", "sched_rt.h"),
                    os.path.join(ROOT, "include", "magma_b200.h")]


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in DEPS):
        return OUT
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wall", "-o", OUT, "-x", "c++", *SCHEDULES, EMUL, EMUL_MODELS]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force=True))
