"""Mesh files on the host side: ctypes face of rayrs_b200/host/mesh_io.{hpp,cpp} — the PLY loader that
completes the reference's commented-out `ply` crate (ply/src/lib.rs) and the OBJ loader of
wavefront_obj.rs:15-44.  Both return the triangle soup that Object::from_triangles (lib.rs:407-415) takes."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _ffi

PLY_FORMATS = {"ascii": 0, "binary_big_endian": 1, "binary_little_endian": 2}


class MeshError(IOError):
    """io::Error (or panic) of the loader; the message is the reference sketch's."""


def _load(path, kind: int) -> np.ndarray:
    lib = _ffi.host_lib()
    n = C.c_uint64(0)
    ptr = lib.rrh_load_mesh(os.fsencode(str(path)), kind, C.byref(n))
    if not ptr:
        raise MeshError(lib.rrh_last_error().decode())
    try:
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(max(n.value, 1) * 9,))[: n.value * 9].copy()
    finally:
        lib.rrh_free(ptr)
    return out.reshape(-1, 3, 3)


def load_ply_file(path) -> np.ndarray:
    """(n, 3, 3) float64 triangles of a PLY file (ascii / binary_little_endian / binary_big_endian; polygons
    are fan-triangulated)."""
    return _load(path, 0)


def load_obj_file(path) -> np.ndarray:
    """wavefront_obj::load_obj_file: 'v x y z' / 'f i j k' lines only."""
    return _load(path, 1)


def load_obj_file_spheres(path) -> np.ndarray:
    """wavefront_obj::load_obj_file_spheres: the (n, 3) vertex positions, one sphere centre each — hand them to
    Object.from_spheres(centers, radius, mat)."""
    lib = _ffi.host_lib()
    n = C.c_uint64(0)
    ptr = lib.rrh_load_obj_spheres(os.fsencode(str(path)), 1.0, C.byref(n))
    if not ptr:
        raise MeshError(lib.rrh_last_error().decode())
    try:
        out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(max(n.value, 1) * 3,))[: n.value * 3].copy()
    finally:
        lib.rrh_free(ptr)
    return out.reshape(-1, 3)


def write_ply(path, vertices: np.ndarray, faces: np.ndarray, fmt: str = "binary_little_endian") -> None:
    """float32 xyz vertices + uchar-count int32 triangle faces."""
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    f = np.ascontiguousarray(faces, dtype=np.int32).reshape(-1, 3)
    lib = _ffi.host_lib()
    if lib.rrh_write_ply(os.fsencode(str(path)), v.ctypes.data, v.shape[0], f.ctypes.data, f.shape[0], PLY_FORMATS[fmt]) != 0:
        raise MeshError(lib.rrh_last_error().decode())


def ply_describe(data: bytes) -> list[str]:
    """Header of an in-memory PLY, one line per item (format / comment / element / property / list);
    parses the body too, so truncated files raise."""
    lib = _ffi.host_lib()
    buf = C.create_string_buffer(1 << 16)
    if lib.rrh_ply_describe(data, len(data), buf, len(buf)) != 0:
        raise MeshError(lib.rrh_last_error().decode())
    return buf.value.decode().splitlines()
