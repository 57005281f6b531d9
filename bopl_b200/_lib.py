"""ctypes binding of libbopl_b200.so — the C-ABI declared in include/bopl_b200.h.

This is the stub a BOPL maintainer would add (INTEGRATION.md shows it next to the reference call sites).
There is no CPU fallback: a missing library raises at load, a missing CUDA device raises at `create()`.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libbopl_b200.so")
SOURCES = ["api.cu", "posterior_trsm_t.cu", "gp_kernels.cu", "uei_kernels.cu", "topk.cu", "fit_kernels.cu", "path_kernels.cu", "posterior_ozaki.cu", "collective.cu", "probes.cu"]
HEADERS = ["common.cuh", "pipeline.cuh", os.path.join("..", "..", "include", "bopl_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-pthread", "-shared", "--threads", "0", "-ldl"]

KERNEL_KINDS = {"Mat52": 0, "rbf": 1, "Mat32": 2, "ExpQuad": 3}
UTILITY_KINDS = {"linear": 0, "quadratic": 1, "exponential": 2, "constrained": 3}
POSTERIOR_KERNELS = {"fp64": 0, "int8": 6, "int8x7": 7, "int8auto": 60}   # bopl_set_posterior_kernel; the library default is int8x7
VAR_INCLUDE_NOISE = 1
VAR_CLIP = 2
ERR_NUMERIC = -5

# every symbol include/bopl_b200.h declares
SYMBOLS = ["bopl_create", "bopl_destroy", "bopl_last_error", "bopl_device_sms", "bopl_build_info", "bopl_set_model",
           "bopl_cross_cov", "bopl_posterior", "bopl_posterior_mean", "bopl_train_mean", "bopl_posterior_grad",
           "bopl_incumbent", "bopl_uei_mc", "bopl_uei", "bopl_uei_grad", "bopl_uei_affine", "bopl_uei_constrained", "bopl_expected_utility", "bopl_topk",
           "bopl_uei_host", "bopl_uei_grad_host", "bopl_launch_count", "bopl_fit_model", "bopl_log_marginal",
           "bopl_get_factor", "bopl_set_posterior_kernel", "bopl_get_posterior_kernel", "bopl_path_begin", "bopl_path_eval_host", "bopl_path_length", "bopl_path_end",
           "bopl_nccl_unique_id", "bopl_nccl_comm_init", "bopl_nccl_comm_destroy", "bopl_topk_allgather", "bopl_topk_allgather_host",
           "bopl_topk_pack", "bopl_topk_merge", "bopl_probe_int8_peak", "bopl_fit_info"]


class BoplError(RuntimeError):
    """A negative bopl_status from the C-ABI (message from bopl_last_error)."""

    def __init__(self, code, message):
        super(BoplError, self).__init__("libbopl_b200 error %d: %s" % (code, message))
        self.code = code


HASH_PATH = LIB_PATH + ".srchash"


def _source_hash():
    """Content hash of every source, header and flag the library is built from (mtimes do not survive a repo snapshot)."""
    import hashlib
    hh = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for name in SOURCES + HEADERS:       # names relative to csrc/: the hash must not depend on where the repo is checked out
        with open(os.path.normpath(os.path.join(CSRC, name)), "rb") as f:
            hh.update(os.path.basename(name).encode() + b"\0" + f.read())
    return hh.hexdigest()


def needs_build():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile libbopl_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    with open(HASH_PATH, "w") as f:
        f.write(_source_hash())
    return LIB_PATH


_lib = None


def load():
    """Load the shared library (raises ImportError if it has not been built — there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("bopl_b200: %s is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    c_int, c_ll, c_dbl, vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_double, ctypes.c_void_p
    ip, dp, llp = vp, vp, vp   # data pointers are passed as raw addresses
    lib.bopl_create.argtypes = [ctypes.POINTER(vp), c_int]
    lib.bopl_destroy.argtypes = [vp]
    lib.bopl_destroy.restype = None
    lib.bopl_last_error.argtypes = [vp]
    lib.bopl_last_error.restype = ctypes.c_char_p
    lib.bopl_device_sms.argtypes = [vp]
    lib.bopl_build_info.argtypes = []
    lib.bopl_build_info.restype = ctypes.c_char_p
    lib.bopl_set_model.argtypes = [vp, c_int, c_int, c_int, c_int, ip, dp, dp, dp, dp, dp, dp, dp, dp, dp]
    lib.bopl_fit_model.argtypes = [vp, c_int, c_int, c_int, c_int, ip, dp, dp, dp, dp, dp, dp, c_int, c_int, dp]
    lib.bopl_fit_info.argtypes = [vp, c_int, ip]
    lib.bopl_log_marginal.argtypes = [vp, c_int, c_int, c_int, ip, dp, dp, dp, dp, dp, dp, dp, ip]
    lib.bopl_get_factor.argtypes = [vp, c_int, c_int, dp, dp, dp, dp]
    lib.bopl_cross_cov.argtypes = [vp, dp, c_int, c_int, c_int, dp, vp]
    lib.bopl_posterior.argtypes = [vp, dp, c_int, c_int, dp, dp, c_int, vp]
    lib.bopl_posterior_mean.argtypes = [vp, dp, c_int, c_int, dp, vp]
    lib.bopl_train_mean.argtypes = [vp, dp, vp]
    lib.bopl_posterior_grad.argtypes = [vp, dp, c_int, c_int, dp, dp, vp]
    lib.bopl_incumbent.argtypes = [vp, c_int, dp, c_int, c_int, dp, vp]
    lib.bopl_uei_mc.argtypes = [vp, dp, dp, c_int, dp, c_int, c_int, dp, c_int, c_int, dp, dp, c_dbl, c_int, dp, vp]
    lib.bopl_uei.argtypes = [vp, dp, c_int, dp, c_int, c_int, dp, c_int, c_int, dp, dp, c_int, dp, vp]
    lib.bopl_uei_grad.argtypes = [vp, dp, c_int, dp, c_int, c_int, dp, c_int, c_int, dp, dp, c_int, dp, dp, vp]
    lib.bopl_uei_affine.argtypes = [vp, dp, c_int, dp, c_int, dp, dp, c_int, dp, dp, vp]
    lib.bopl_expected_utility.argtypes = [vp, dp, c_int, dp, c_int, c_int, dp, c_int, c_int, dp, c_int, dp, dp, vp]
    lib.bopl_uei_constrained.argtypes = [vp, dp, c_int, dp, c_int, dp, dp, c_int, dp, dp, vp]
    lib.bopl_topk.argtypes = [vp, dp, c_int, c_int, c_ll, dp, llp, vp]
    lib.bopl_uei_host.argtypes = [vp, dp, c_int, dp, c_int, c_int, dp, c_int, c_int, dp, dp, c_int, dp, c_int, c_ll,
                                  dp, llp]
    lib.bopl_uei_grad_host.argtypes = [vp, dp, c_int, dp, c_int, c_int, dp, c_int, c_int, dp, dp, c_int, dp, dp]
    lib.bopl_path_begin.argtypes = [vp, c_int, c_int, dp]
    lib.bopl_path_eval_host.argtypes = [vp, dp, dp, dp, dp, dp]
    lib.bopl_path_length.argtypes = [vp]
    lib.bopl_path_end.argtypes = [vp]
    lib.bopl_launch_count.argtypes = [vp]
    lib.bopl_probe_int8_peak.argtypes = [vp, c_dbl, ctypes.POINTER(c_dbl), ctypes.POINTER(c_dbl)]
    lib.bopl_nccl_unique_id.argtypes = [vp]
    lib.bopl_nccl_comm_init.argtypes = [vp, c_int, c_int, vp, ctypes.POINTER(vp)]
    lib.bopl_nccl_comm_destroy.argtypes = [vp]
    lib.bopl_topk_allgather.argtypes = [vp, vp, dp, llp, c_int, dp, c_ll, c_int, dp, llp, dp, vp]
    lib.bopl_topk_allgather_host.argtypes = [vp, vp, dp, llp, c_int, dp, c_int, dp, llp, dp]
    lib.bopl_topk_pack.argtypes = [vp, dp, llp, c_int, dp, c_ll, c_int, dp, vp]
    lib.bopl_topk_merge.argtypes = [vp, dp, c_int, c_int, dp, llp, dp, vp]
    lib.bopl_set_posterior_kernel.argtypes = [vp, c_int]
    lib.bopl_get_posterior_kernel.argtypes = [vp]
    lib.bopl_launch_count.restype = c_ll
    for name in SYMBOLS:
        getattr(lib, name)   # AttributeError if the library does not export what the header declares
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != 0:
        msg = load().bopl_last_error(handle)
        raise BoplError(rc, msg.decode("utf-8", "replace") if msg else "")
