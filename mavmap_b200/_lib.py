"""Loader for the CUDA library (mavmap_b200/libmavmap_b200.so).

There is no CPU fallback: if the library is missing, or no CUDA device is usable,
every compute call raises.  The .so is built in-tree by __graft_entry__.build().
"""
import ctypes as C
import os

from . import _abi

_LIB = None
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmavmap_b200.so")


class MavmapB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mavmap_b200 error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        _LIB = _abi.bind(C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL))
        if _LIB.mm_abi_version() != 1:
            raise ImportError("ABI version mismatch in %s" % LIB_PATH)
    return _LIB


def check(code):
    if code == _abi.MM_OK:
        return
    msg = lib().mm_last_error()
    msg = msg.decode() if msg else ""
    if code == _abi.MM_ERR_DATUM or code == _abi.MM_ERR_MIN_TRACK_LEN or code == _abi.MM_ERR_INVALID_ARG:
        raise ValueError(msg or "invalid argument")      # std::invalid_argument analogue
    raise MavmapB200Error(code, msg)


def device_count():
    return lib().mm_device_count()


def kernel_launch_count():
    return int(lib().mm_kernel_launch_count())


def match_pair_counters():
    """(calls, descriptor arrays uploaded, bytes uploaded) of mm_match_pair since process start."""
    import ctypes as C
    a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    lib().mm_match_pair_counters(C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value
