import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run on the B200 box)")


def _have_gpu():
    try:
        import ctypes
        cuda = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cuda.cuInit(0) != 0:
            return False
        cuda.cuDeviceGetCount(ctypes.byref(n))
        return n.value > 0
    except OSError:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_test_infrastructure():
    """oracle/ (gcc) and the kernel emulation harness (g++).  The product library is built by
    __graft_entry__.build(); tests that need it fail loudly if it is absent."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    emu_dir = os.path.join(ROOT, "tests", "emu")
    so = os.path.join(emu_dir, "libemu_kernels.so")
    src = os.path.join(emu_dir, "emu_kernels.cpp")
    deps = [src, os.path.join(emu_dir, "cuda_emu.h")]
    csrc = os.path.join(ROOT, "stm32f7-rtlsdr_b200", "csrc")
    deps += [os.path.join(csrc, f) for f in os.listdir(csrc)]
    deps.append(os.path.join(ROOT, "include", "b200sdr_synth.h"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    yield


@pytest.fixture(scope="session")
def sdr_lib():
    """The product library, built if needed (nvcc cross-compiles without a GPU)."""
    import importlib
    pkg = importlib.import_module("stm32f7-rtlsdr_b200")
    build = importlib.import_module("stm32f7-rtlsdr_b200.build")
    build.build()
    return pkg
