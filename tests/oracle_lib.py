"""TEST INFRASTRUCTURE: loader for the CPU oracle (oracle/libra_oracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may use this; the product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from ra_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libra_oracle.so")
_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        src = os.path.join(_ROOT, "oracle", "ra_oracle.c")
        if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_SO)
        _lib.ra_oracle_agreed_commit.restype = C.c_uint64
        _lib.ra_oracle_agreed_commit.argtypes = [C.POINTER(C.c_uint64), C.c_size_t]
        _lib.ra_oracle_flood.restype = C.c_int
        _lib.ra_oracle_flood.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64,
                                         C.c_uint32]
    return _lib


class Oracle(abi.Backend):
    name = "oracle"

    def __init__(self, n_groups: int, n_members: int, **kw):
        super().__init__(lib(), "ra_oracle", n_groups, n_members, **kw)

    def flood(self, n_steps: int, cmds_per_step: int = 1, election_permille: int = 0, seed: int = 1,
              threads: int = 1, faults=None) -> None:
        if faults is None:
            self._check(lib().ra_oracle_flood(self._h, n_steps, cmds_per_step, election_permille, seed,
                                              threads), "flood")
            return
        f = lib().ra_oracle_flood_faults
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.POINTER(abi.RaFloodFaults)]
        ff = abi.RaFloodFaults(*faults)
        self._check(f(self._h, n_steps, cmds_per_step, election_permille, seed, threads, C.byref(ff)), "flood_faults")


    def set_sample(self, stride: int, offset: int, total_groups: int) -> None:
        """group g of this oracle plays global group offset + g * stride of a flood over total_groups groups"""
        f = lib().ra_oracle_set_sample
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        self._check(f(self._h, stride, offset, total_groups), "set_sample")


def agreed_commit(indexes):
    arr = (C.c_uint64 * len(indexes))(*indexes)
    return int(lib().ra_oracle_agreed_commit(arr, len(indexes)))
