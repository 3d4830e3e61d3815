"""ctypes binding of libthejoker_b200.so (include/thejoker_b200.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a) and loaded from the
package directory.  There is no fallback: if the shared object is missing or CUDA is
unusable the product raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TJB_LIB_PATH", os.path.join(_HERE, "libthejoker_b200.so"))  # override: tuning builds
_SRC = [os.path.join(_HERE, "csrc", f) for f in
        ("tjb_api.cu", "kepler.cuh", "linalg.cuh", "marginal_ll.cuh", "accept.cuh", "posterior.cuh",
         "prior_gen.cuh", "comm.hpp", "star_tables.hpp")]
_HDR = os.path.join(os.path.dirname(_HERE), "include", "thejoker_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

TJB_MAX_LINEAR = 8
TJB_COMM_ID_BYTES = 128
_dp = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)


class TjbSpec(ctypes.Structure):
    _fields_ = [
        ("n_times", ctypes.c_int32),
        ("n_linear", ctypes.c_int32),
        ("t_ref", ctypes.c_double),
        ("t", _dp),
        ("rv", _dp),
        ("ivar", _dp),
        ("trend_M", _dp),
        ("mu", ctypes.c_double * TJB_MAX_LINEAR),
        ("Lambda", ctypes.c_double * TJB_MAX_LINEAR),
        ("K_prior_kind", ctypes.c_int32),
        ("sigma_K0", ctypes.c_double),
        ("P0", ctypes.c_double),
        ("max_K", ctypes.c_double),
        ("jitter_mode", ctypes.c_int32),
    ]


class TjbPcg64(ctypes.Structure):
    _fields_ = [("state_hi", ctypes.c_uint64), ("state_lo", ctypes.c_uint64),
                ("inc_hi", ctypes.c_uint64), ("inc_lo", ctypes.c_uint64)]


class TjbPriorDist(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("reserved", ctypes.c_int32), ("p0", ctypes.c_double),
                ("p1", ctypes.c_double), ("scale", ctypes.c_double)]


class TjbPriorGen(ctypes.Structure):
    _fields_ = [("par", TjbPriorDist * 5), ("seed", ctypes.c_uint64)]


class TjbMultiStarJob(ctypes.Structure):
    _fields_ = [
        ("n_stars", ctypes.c_int64),
        ("specs", ctypes.POINTER(TjbSpec)),
        ("pcg", ctypes.POINTER(TjbPcg64)),
        ("d_P", ctypes.c_void_p), ("d_e", ctypes.c_void_p), ("d_omega", ctypes.c_void_p),
        ("d_M0", ctypes.c_void_p), ("d_s", ctypes.c_void_p),
        ("s_const", ctypes.c_double),
        ("n_prior", ctypes.c_int64),
        ("max_keep", ctypes.c_int64),
        ("near_tol", ctypes.c_double),
        ("n_per", ctypes.c_int32),
        ("clamp_K", ctypes.c_int32),
        ("n_slots", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
        ("h_normals", ctypes.c_void_p),
        ("h_idx", ctypes.c_void_p),
        ("h_counts", ctypes.c_void_p),
        ("h_llmax", ctypes.c_void_p),
        ("h_rows", ctypes.c_void_p),
        ("h_ll", ctypes.c_void_p),
    ]


# every symbol include/thejoker_b200.h declares: name -> (restype, argtypes)
_vp = ctypes.c_void_p
_H = ctypes.c_void_p  # TjbHandle*
SYMBOLS = {
    "tjb_create": (ctypes.c_int, [ctypes.POINTER(TjbSpec), ctypes.c_int, ctypes.POINTER(_H)]),
    "tjb_update_star": (ctypes.c_int, [_H, ctypes.POINTER(TjbSpec)]),
    "tjb_destroy": (None, [_H]),
    "tjb_set_stream": (ctypes.c_int, [_H, _vp]),
    "tjb_last_error": (ctypes.c_char_p, []),
    "tjb_version": (ctypes.c_int, []),
    "tjb_device_info": (ctypes.c_int, [_H] + [ctypes.POINTER(ctypes.c_int)] * 4),
    "tjb_marginal_ll_soa": (ctypes.c_int, [_H, _vp, _vp, _vp, _vp, _vp, ctypes.c_double,
                                           ctypes.c_int64, _vp, _vp]),
    "tjb_set_peer_keys": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_void_p),
                                         ctypes.POINTER(ctypes.c_int), ctypes.c_int]),
    "tjb_marginal_ll_aos": (ctypes.c_int, [_H, _vp, ctypes.c_int, ctypes.c_int64, _vp, _vp]),
    "tjb_marginal_ll_host": (ctypes.c_int, [_H, _vp, ctypes.c_int64, _vp]),
    "tjb_marginal_ll_host_soa": (ctypes.c_int, [_H, _vp, _vp, _vp, _vp, _vp, ctypes.c_double,
                                                ctypes.c_int64, _vp]),
    "tjb_marginal_ll_host_soa_resident": (ctypes.c_int, [_H, _vp, _vp, _vp, _vp, _vp,
                                                         ctypes.c_double, ctypes.c_int64, _vp,
                                                         _vp]),
    "tjb_prior_sample": (ctypes.c_int, [ctypes.c_int, _vp, ctypes.POINTER(TjbPriorGen),
                                        ctypes.c_int64, ctypes.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "tjb_prior_rows": (ctypes.c_int, [_H, ctypes.POINTER(TjbPriorGen), _vp, ctypes.c_int64, _vp]),
    "tjb_marginal_ll_generated": (ctypes.c_int, [_H, ctypes.POINTER(TjbPriorGen), ctypes.c_int64,
                                                 ctypes.c_int64, _vp, _vp]),
    "tjb_llmax_reset": (ctypes.c_int, [_H, _vp]),
    "tjb_llmax_update": (ctypes.c_int, [_H, _vp, ctypes.c_int64, _vp]),
    "tjb_llmax_get": (ctypes.c_int, [_H, _vp, _dp]),
    "tjb_key_to_double": (ctypes.c_double, [ctypes.c_int64]),
    "tjb_double_to_key": (ctypes.c_int64, [ctypes.c_double]),
    "tjb_accept": (ctypes.c_int, [_H, _vp, ctypes.c_int64, _vp, _vp, ctypes.POINTER(TjbPcg64),
                                  ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
                                  _vp, _i64p]),
    "tjb_accept_nonfinite": (ctypes.c_int, [_H, _i64p]),
    "tjb_comm_unique_id": (ctypes.c_int, [_vp]),
    "tjb_comm_create": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_void_p)]),
    "tjb_comm_destroy": (None, [_vp]),
    "tjb_comm_allreduce_max_key": (ctypes.c_int, [_H, _vp, _vp]),
    "tjb_accept_dist": (ctypes.c_int, [_H, _vp, _vp, ctypes.c_int64, _vp, _vp,
                                       ctypes.POINTER(TjbPcg64), ctypes.c_int64, ctypes.c_int64,
                                       ctypes.c_double, _vp, _i64p]),
    "tjb_pcg64_uniform": (ctypes.c_int, [_H, ctypes.POINTER(TjbPcg64), ctypes.c_int64,
                                         ctypes.c_int64, _vp]),
    "tjb_posterior_aA": (ctypes.c_int, [_H, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp]),
    "tjb_posterior_draw": (ctypes.c_int, [_H, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp,
                                          _vp, _vp]),
    "tjb_multistar_rejection": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(TjbMultiStarJob)]),
    "tjb_unmarginalized_ll": (ctypes.c_int, [_H, _vp, ctypes.c_int64, _vp]),
    "tjb_design_column": (ctypes.c_int, [_H, _vp, _vp, _vp]),
    "tjb_get_stats": (ctypes.c_int, [_H, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int]),
    "tjb_set_epoch_rows_mode": (ctypes.c_int, [ctypes.c_int]),
    "tjb_fp64_peak": (ctypes.c_int, [_H, ctypes.c_int, _dp, _dp]),
}

_lib = None


_KERNEL_SRC = ("kepler.cuh", "linalg.cuh", "marginal_ll.cuh", "prior_gen.cuh")


def source_hash() -> str:
    """sha256 over the sources the likelihood kernel is compiled from (kepler.cuh,
    linalg.cuh, marginal_ll.cuh, prior_gen.cuh) and the nvcc flags: identifies the kernel
    code a built library -- and an ncu profile of it -- belongs to.
    profiles/kernel_counts.json carries the hash its counters were measured on; bench.py
    flags counters whose hash differs from the sources in the tree."""
    import hashlib

    h = hashlib.sha256()
    for name in _KERNEL_SRC:
        h.update(name.encode())
        with open(os.path.join(_HERE, "csrc", name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    mt = os.path.getmtime(LIB_PATH)
    return any(os.path.exists(s) and os.path.getmtime(s) > mt for s in _SRC + [_HDR])


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    if force or needs_build():
        cmd = ["nvcc"] + NVCC_FLAGS + [_SRC[0], "-o", LIB_PATH]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout)
        if verbose:
            print(res.stdout)
    return LIB_PATH


def load():
    """dlopen the library and declare every entry point.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; "
                           "g.build()'` (nvcc, sm_100a).  thejoker_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class TjbError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        msg = load().tjb_last_error()
        raise TjbError(f"libthejoker_b200 error {rc}: {msg.decode() if msg else ''}")
