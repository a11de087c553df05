"""TEST INFRASTRUCTURE ONLY: loads the CPU-emulated build of the kernel sources (tests/cpu_emu)."""
import ctypes
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_EMU = None


def emu_lib():
    global _EMU
    if _EMU is None:
        spec = importlib.util.spec_from_file_location("_build_emu", os.path.join(_HERE, "cpu_emu", "build_emu.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        path = mod.build()
        from deformationpyramid_b200 import _lib
        lib = _lib.bind(ctypes.CDLL(path))
        lib._ndp_requires_cuda = False
        lib.ndp_hook_point_forward.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2 + [ctypes.c_longlong] + [ctypes.c_void_p] * 2
        lib.ndp_hook_point_backward.argtypes = [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4 + [ctypes.c_longlong] + [ctypes.c_void_p] * 2
        lib.ndp_hook_head_dim.argtypes = [ctypes.c_int] * 3
        _EMU = lib
    return _EMU
