"""ctypes binding of liboard_b200.so (C ABI in include/oard.h).

The shared library is built in-tree by `build()` (nvcc, sm_100a) — see __graft_entry__.build().  There is no CPU
fallback: importing this module without the library, or calling compute entry points without a CUDA device, raises.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboard_b200.so")
SOURCES = [os.path.join(_HERE, "csrc", "oard.cu")]
HEADER = os.path.join(os.path.dirname(_HERE), "include", "oard.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-shared", "-Xcompiler", "-fPIC"]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu -> liboard_b200.so for sm_100a (cross-compiles without a GPU)."""
    deps = SOURCES + [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))] + [HEADER]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


class OardCfg(C.Structure):
    _fields_ = [("hidden_channels", C.c_int32), ("num_radial", C.c_int32), ("num_layers", C.c_int32),
                ("in_hidden_channels", C.c_int32), ("cutoff", C.c_float), ("reflect_equiv", C.c_int32),
                ("legacy", C.c_int32), ("update", C.c_int32), ("object_aware", C.c_int32)]


# every symbol include/oard.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "oard_abi_version": (C.c_int, []),
    "oard_last_error": (C.c_char_p, []),
    "oard_create": (C.c_int, [C.POINTER(OardCfg), C.c_int, C.POINTER(C.c_void_p)]),
    "oard_destroy": (None, [C.c_void_p]),
    "oard_num_weights": (C.c_int, [C.c_void_p]),
    "oard_weight_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "oard_weight_numel": (C.c_int64, [C.c_void_p, C.c_int]),
    "oard_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "oard_commit_weights": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oard_plan": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "oard_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "oard_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oard_dyn_configure": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "oard_dyn_num_weights": (C.c_int, [C.c_void_p]),
    "oard_dyn_weight_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "oard_dyn_weight_numel": (C.c_int64, [C.c_void_p, C.c_int]),
    "oard_dyn_set_weight": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    "oard_dyn_plan": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "oard_dyn_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oard_reverse_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_void_p]),
    "oard_inpaint_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_float, C.c_void_p]),
    "oard_jump_back": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p]),
    "oard_forward_train": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oard_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "oard_zero_grads": (C.c_int, [C.c_void_p, C.c_void_p]),
    "oard_get_grad": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "oard_set_debug": (C.c_int, [C.c_void_p, C.c_int]),
    "oard_debug_bytes": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "oard_debug_read": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
    "oard_last_launch_count": (C.c_int64, [C.c_void_p]),
    "oard_total_launch_count": (C.c_int64, [C.c_void_p]),
    "oard_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "oard_profile_count": (C.c_int, [C.c_void_p]),
    "oard_test_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "oard_test_gemm_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                    C.POINTER(C.c_float), C.c_void_p]),
    "oard_test_gemm_p16": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                     C.POINTER(C.c_float), C.c_void_p]),
    "oard_profile_get": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_double),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load the library (once).  Raises if it has not been built: the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). oareactdiff_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class OardError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise OardError(f"oard error {rc}: {load().oard_last_error().decode()}")
