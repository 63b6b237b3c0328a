"""Build and load the SIMT-emulated build of the product sources (tests only)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_SO = os.path.join(EMU_DIR, "_build", "libsmallz4_emu.so")
CSRC = os.path.join(ROOT, "smallz4_b200", "csrc")


def build_emu():
    os.makedirs(os.path.dirname(EMU_SO), exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(EMU_DIR, f) for f in os.listdir(EMU_DIR)]
    newest = max(os.path.getmtime(s) for s in srcs if os.path.isfile(s))
    if os.path.exists(EMU_SO) and os.path.getmtime(EMU_SO) >= newest:
        return EMU_SO
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-DSZ4_EMU", "-DSZ4_DP_SEG=4096", "-DSZ4_DP_WARM=512", "-DSZ4_DP_SLACK=128", "-DSZ4_DP_RING=512", "-DSZ4_GREEDY_SEG=4096", "-DSZ4_GREEDY_WARM=256", "-DSZ4_PATH_SEG=4096", "-DSZ4_PATH_WARM=256", "-DSZ4_LSD_CHUNK=65536", "-I" + EMU_DIR, "-I" + CSRC, "-x", "c++",
                           os.path.join(CSRC, "sz4_pipeline.cu"), "-x", "c++", os.path.join(EMU_DIR, "cuda_emu.cpp"),
                           "-shared", "-fPIC", "-o", EMU_SO])
    return EMU_SO


def emu_compressor(**options):
    from smallz4_b200.api import Compressor
    return Compressor(lib_path=build_emu(), **options)
