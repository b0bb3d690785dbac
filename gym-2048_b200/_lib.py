"""ctypes binding of libg2048.so (include/g2048.h) and its in-tree nvcc build.

There is no CPU implementation behind this module: if the shared library is missing and
cannot be built, or a call fails, an exception is raised.
"""
import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
SO_PATH = os.environ.get("G2048_SO") or os.path.join(_PKG, "libg2048.so")     # G2048_SO: a variant build (experiments)
SOURCES = [os.path.join(_PKG, "csrc", f) for f in ("g2048.cu", "g2048_data.cu", "g2048_csv.cpp")]
HEADERS = [os.path.join(_PKG, "csrc", "g2048_device.cuh"), os.path.join(_PKG, "csrc", "g2048_internal.h"),
           os.path.join(_ROOT, "include", "g2048.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]

ABI_VERSION = 5
FLAG_AUTO_RESET = 1
FLAG_POLICY_UNIFORM, FLAG_POLICY_LEGAL = 2, 4
FLAG_CHAIN_INTERLEAVED = 16  # G2048_FLAG_CHAIN_INTERLEAVED: launches of other chains in between (launch shape)
FLAG_CHAINED = 8            # G2048_FLAG_CHAINED: per-slice dependency on the previous launch of the same chain buffer
CHAIN_WORDS = 16384          # G2048_CHAIN_WORDS (64-bit words of a chain buffer)
OBS_U8, OBS_F32, OBS_I64, OBS_BF16 = 0, 1, 2, 3

EXPORTS = [
    "g2048_abi_version", "g2048_last_error", "g2048_step", "g2048_step_n", "g2048_step_list", "g2048_step_list_timed", "g2048_one", "g2048_step_many", "g2048_reset", "g2048_add_tile", "g2048_move", "g2048_status",
    "g2048_encode_obs", "g2048_values_from_exp", "g2048_exp_from_values", "g2048_philox", "g2048_philox2x32", "g2048_draw_words",
    "g2048_env_create", "g2048_env_destroy", "g2048_env_reset_host", "g2048_env_step_host",
    "g2048_env_device_ptrs", "g2048_env_set_boards_host", "g2048_env_get_boards_host", "g2048_env_step_index",
    "g2048_unpack_boards_host",
    "g2048_sample_actions", "g2048_symmetry", "g2048_augment", "g2048_discounted_return", "g2048_gae",
    "g2048_csv_export", "g2048_csv_rows", "g2048_csv_import",
]


class G2048Error(RuntimeError):
    pass


class StepArgs(C.Structure):
    """G2048StepArgs (include/g2048.h)."""
    _fields_ = [
        ("boards", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p),
        ("dones", C.c_void_p), ("illegal", C.c_void_p), ("highest_exp", C.c_void_p),
        ("legal_mask", C.c_void_p), ("terminal_boards", C.c_void_p),
        ("ep_score", C.c_void_p), ("ep_len", C.c_void_p),
        ("final_score", C.c_void_p), ("final_len", C.c_void_p),
        ("forced_draws", C.c_void_p), ("step_counter", C.c_void_p),
        ("n", C.c_uint64), ("env_id_base", C.c_uint64), ("seed", C.c_uint64),
        ("step_index", C.c_uint64),
        ("illegal_move_reward", C.c_float), ("max_tile_exp", C.c_uint32),
        ("flags", C.c_uint32), ("boards_out", C.c_void_p),
        ("ep_return", C.c_void_p), ("final_return", C.c_void_p),
        ("boards_nibble", C.c_void_p), ("nibble_overflow", C.c_void_p), ("chain", C.c_void_p),
    ]


class OneIO(C.Structure):
    """G2048OneIO (include/g2048.h): the packed in/out block of g2048_one."""
    _fields_ = [
        ("values", C.c_int64 * 16), ("obs", C.c_int64 * 256), ("reward", C.c_float), ("score", C.c_uint32),
        ("done", C.c_uint8), ("illegal", C.c_uint8), ("changed", C.c_uint8), ("highest_exp", C.c_uint8),
        ("legal_mask", C.c_uint8), ("n_empty", C.c_uint8), ("is_end", C.c_uint8), ("bad_cells", C.c_uint8),
    ]


ONE_STEP, ONE_RESET, ONE_MOVE, ONE_ADD_TILE, ONE_STATUS = 0, 1, 2, 3, 4


class StepManyArgs(C.Structure):
    """G2048StepManyArgs (include/g2048.h)."""
    _fields_ = [
        ("boards", C.c_void_p), ("actions", C.c_void_p), ("rewards", C.c_void_p), ("dones", C.c_void_p),
        ("illegal", C.c_void_p), ("boards_traj", C.c_void_p),
        ("n", C.c_uint64), ("env_id_base", C.c_uint64), ("seed", C.c_uint64), ("step_index", C.c_uint64),
        ("n_steps", C.c_uint32), ("illegal_move_reward", C.c_float), ("max_tile_exp", C.c_uint32),
        ("flags", C.c_uint32), ("actions_out", C.c_void_p), ("legal_mask", C.c_void_p),
    ]


class EnvConfig(C.Structure):
    """G2048EnvConfig (include/g2048.h)."""
    _fields_ = [
        ("device", C.c_int32), ("flags", C.c_uint32), ("n", C.c_uint64),
        ("env_id_base", C.c_uint64), ("seed", C.c_uint64),
        ("illegal_move_reward", C.c_float), ("max_tile_exp", C.c_uint32),
        ("n_chunks", C.c_uint32), ("board_format", C.c_uint32), ("unpack_threads", C.c_uint32),
    ]


class HostStepOut(C.Structure):
    """G2048HostStepOut (include/g2048.h)."""
    _fields_ = [
        ("boards", C.c_void_p), ("rewards", C.c_void_p), ("dones", C.c_void_p),
        ("illegal", C.c_void_p), ("highest_exp", C.c_void_p), ("legal_mask", C.c_void_p),
        ("nibble_overflow", C.c_void_p),
    ]


BOARDS_BYTES, BOARDS_NIBBLE, BOARDS_BYTES_PACKED_WIRE = 0, 1, 2


def _stale():
    if os.environ.get("G2048_SO"):
        return False                    # an explicitly chosen library is used as it is
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    return any(os.path.exists(f) and os.path.getmtime(f) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile csrc/g2048.cu for sm_100a into gym-2048_b200/libg2048.so (nvcc cross-compiles
    without a GPU).  Returns the path.

    Safe under torchrun on a fresh checkout: the build runs under an exclusive file lock (the ranks that
    lose the race wait and find the library up to date), nvcc links into a temporary file in the same
    directory and the result is moved into place atomically, so no process can dlopen a half-written file."""
    if not force and not _stale():
        return SO_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise G2048Error("nvcc not found: cannot build %s" % SO_PATH)
    import fcntl
    with open(SO_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not _stale():          # another process built it while we waited
                return SO_PATH
            tmp = "%s.tmp.%d" % (SO_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise G2048Error("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
            os.replace(tmp, SO_PATH)
            if verbose:
                print(r.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return SO_PATH


_lib = None


def lib():
    """The loaded library.  Raises G2048Error when it is absent and cannot be built."""
    global _lib
    if _lib is not None:
        return _lib
    if _stale():
        try:
            build()
        except G2048Error:
            if not os.path.exists(SO_PATH):
                raise
    try:
        L = C.CDLL(SO_PATH)
    except OSError as e:
        raise G2048Error("cannot load %s: %s (no CPU fallback exists)" % (SO_PATH, e))
    L.g2048_last_error.restype = C.c_char_p
    L.g2048_env_step_index.restype = C.c_uint64
    u64, vp, u32 = C.c_uint64, C.c_void_p, C.c_uint32
    L.g2048_step.argtypes = [C.POINTER(StepArgs), vp]
    L.g2048_step_n.argtypes = [C.POINTER(StepArgs), u32, u64, vp]
    L.g2048_step_list.argtypes = [vp, u64, vp]
    L.g2048_step_list_timed.argtypes = [vp, u64, u64, vp, u64, vp]
    L.g2048_one.argtypes = [vp, C.c_int, C.c_int, C.c_int, u64, u64, C.c_float, u32, vp, C.c_int]
    L.g2048_step_many.argtypes = [C.POINTER(StepManyArgs), vp]
    L.g2048_reset.argtypes = [vp, vp, u64, u64, u64, u64, vp]
    L.g2048_add_tile.argtypes = [vp, u64, u64, u64, u64, vp]
    L.g2048_move.argtypes = [vp, vp, vp, vp, vp, u64, vp]
    L.g2048_status.argtypes = [vp, vp, vp, vp, vp, u32, u64, vp]
    L.g2048_encode_obs.argtypes = [vp, vp, C.c_int, u64, vp]
    L.g2048_values_from_exp.argtypes = [vp, vp, u64, vp]
    L.g2048_exp_from_values.argtypes = [vp, vp, u64, vp, vp]
    L.g2048_philox.argtypes = [vp, u32, u32, vp, u64, vp]
    L.g2048_philox2x32.argtypes = [vp, u32, vp, u64, vp]
    L.g2048_draw_words.argtypes = [vp, u64, u64, u64, u64, u32, vp]
    L.g2048_env_create.argtypes = [C.POINTER(vp), C.POINTER(EnvConfig)]
    L.g2048_env_destroy.argtypes = [vp]
    L.g2048_env_reset_host.argtypes = [vp, vp]
    L.g2048_env_step_host.argtypes = [vp, vp, C.POINTER(HostStepOut)]
    L.g2048_env_device_ptrs.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.g2048_env_set_boards_host.argtypes = [vp, vp]
    L.g2048_env_get_boards_host.argtypes = [vp, vp]
    L.g2048_env_step_index.argtypes = [vp]
    L.g2048_unpack_boards_host.argtypes = [vp, vp, C.c_uint64]
    L.g2048_sample_actions.argtypes = [vp, vp, u64, u64, u64, u64, vp]
    L.g2048_symmetry.argtypes = [vp, vp, vp, vp, vp, vp, u64, C.c_int, C.c_int, vp]
    L.g2048_augment.argtypes = [vp] * 5 + [u64] + [vp] * 6
    L.g2048_discounted_return.argtypes = [vp, vp, vp, u64, C.c_double, vp]
    L.g2048_gae.argtypes = [vp] * 7 + [u64, u64, C.c_double, C.c_double, vp]
    L.g2048_csv_export.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, vp, u64, C.c_int]
    L.g2048_csv_rows.argtypes = [C.c_char_p, C.POINTER(u64), C.POINTER(C.c_int)]
    L.g2048_csv_import.argtypes = [C.c_char_p, vp, vp, vp, vp, vp, vp, u64]
    if L.g2048_abi_version() != ABI_VERSION:
        raise G2048Error("libg2048.so ABI %d != binding ABI %d" % (L.g2048_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise G2048Error("g2048 error %d: %s" % (rc, lib().g2048_last_error().decode()))
