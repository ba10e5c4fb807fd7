"""ctypes binding of include/cabana_b200.h (the C-ABI drop-in boundary).

The library is the product: if it is missing this module raises -- there is no CPU
fallback anywhere in cabana_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libcabana_b200.so")
HEADER_PATH = os.path.join(ROOT, "include", "cabana_b200.h")

CB_OK, CB_ERR_INVALID, CB_ERR_CUDA, CB_ERR_OVERFLOW, CB_ERR_UNSUPPORTED, CB_ERR_NOMEM = range(6)
FULL, HALF = 0, 1
CSR, LAYOUT_2D = 0, 1
OP_SERIAL, OP_TEAM, OP_TEAM_VECTOR = 0, 1, 2
ROWS_REFERENCE, ROWS_BINNED = 0, 1


class Positions(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("n", C.c_int64),
        ("outer_stride", C.c_int64),
        ("vlen", C.c_int32),
        ("comp_stride", C.c_int64),
    ]


class Field(C.Structure):
    _fields_ = [
        ("base", C.c_void_p),
        ("n", C.c_int64),
        ("outer_stride", C.c_int64),
        ("vlen", C.c_int32),
        ("comp_stride", C.c_int64),
        ("num_comp", C.c_int32),
        ("elem_bytes", C.c_int32),
    ]


class Grid(C.Structure):
    _fields_ = [
        ("min", C.c_double * 3),
        ("max", C.c_double * 3),
        ("dx", C.c_double * 3),
        ("rdx", C.c_double * 3),
        ("nx", C.c_int32 * 3),
    ]


class LclView(C.Structure):
    _fields_ = [
        ("grid", Grid),
        ("stencil_grid", Grid),
        ("cell_range", C.c_int32),
        ("sorted", C.c_int32),
        ("begin", C.c_int64),
        ("end", C.c_int64),
        ("num_cells", C.c_int64),
        ("counts", C.c_void_p),
        ("offsets", C.c_void_p),
        ("permute", C.c_void_p),
        ("particle_bins", C.c_void_p),
    ]


class VerletView(C.Structure):
    _fields_ = [
        ("layout", C.c_int32),
        ("algorithm", C.c_int32),
        ("n", C.c_int64),
        ("counts", C.c_void_p),
        ("offsets", C.c_void_p),
        ("neighbors", C.c_void_p),
        ("total", C.c_int64),
        ("max_n", C.c_int64),
        ("width", C.c_int64),
        ("row_stride", C.c_int64),
        ("col_stride", C.c_int64),
        ("refilled", C.c_int32),
        ("extent", C.c_int64),
    ]


class CabanaB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"cabana_b200 error {code}: {msg}")
        self.code = code


_lib = None


def declared_symbols() -> list[str]:
    """Every function include/cabana_b200.h declares (used by the CPU symbol test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CabanaB200Error(
            -1,
            f"{LIB_PATH} is missing: build it with `python -m cabana_b200.build` "
            "(the CUDA extension is the product; there is no CPU fallback)",
        )
    L = C.CDLL(LIB_PATH)
    L.cb_last_error_string.restype = C.c_char_p
    L.cb_grid_min_distance_to_point.restype = C.c_double
    L.cb_comm_tuple_bytes.restype = C.c_int64
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != CB_OK:
        raise CabanaB200Error(rc, lib().cb_last_error_string().decode())


def d3(v):
    return (C.c_double * 3)(*[float(a) for a in v])
