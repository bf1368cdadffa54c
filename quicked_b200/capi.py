"""ctypes bindings of the C-ABI declared in include/quicked.h and include/quicked_b200.h.

This is the Python host-side mirror of the reference's operator interface:
  * `QuickedAligner` has the method surface of the reference's C++/pybind11 class
    (reference bindings/cpp/quicked.hpp:46-73, bindings/python/quicked.cpp:30-64): align, set*, getScore, getCigar;
  * `BatchAligner` is the additive batched entry point (one GPU context, packed batch in, scores + CIGARs out).

The product path has no CPU fallback: loading fails loudly when libquicked_b200.so is missing, and every compute
call raises when there is no CUDA device.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libquicked_b200.so")

QUICKED, WINDOWED, BANDED, HIRSCHBERG = 0, 1, 2, 3
QUICKED_OK, QUICKED_ERROR, QUICKED_FAIL_NON_CONVERGENCE = 0, -1, -2
QUICKED_UNKNOWN_ALGO, QUICKED_EMPTY_SEQUENCE, QUICKED_UNIMPLEMENTED, QUICKED_WIP = -3, -4, -10, 1
QB200_ERR_NO_DEVICE, QB200_ERR_CUDA, QB200_ERR_ARG, QB200_ERR_OOM, QB200_ERR_CAPACITY = -100, -101, -102, -103, -104

# every symbol the two headers declare (tests/test_cabi.py checks the library exports all of them)
EXPORTS = ["quicked_check_error", "quicked_status_msg", "quicked_default_params", "quicked_new", "quicked_free",
           "quicked_align", "qb200_device_count", "qb200_create", "qb200_destroy", "qb200_set_stream",
           "qb200_set_workspace_limit", "qb200_last_error", "qb200_upload", "qb200_upload_device", "qb200_run",
           "qb200_download", "qb200_get_stats", "qb200_get_bounds", "qb200_cigar_to_sam", "qb200_align_batch", "qb200_host_alloc", "qb200_host_free",
           "qb200_generate_pairs", "qb200_generate_pairs_ex", "qb200_measure_int_peak", "qb200_pack_batch", "qb200_upload_packed",
           "qb200_align_batch_packed", "qb200_generate_device", "qb200_batch_bytes", "qb200_download_batch"]


class Params(C.Structure):        # quicked_params_t, 48 bytes
    _fields_ = [("algo", C.c_int), ("bandwidth", C.c_uint), ("window_size", C.c_uint), ("overlap_size", C.c_uint),
                ("hew_threshold", C.c_uint * 2), ("hew_percentage", C.c_uint * 2), ("only_score", C.c_bool),
                ("force_scalar", C.c_bool), ("external_timer", C.c_bool), ("external_allocator", C.c_void_p)]


class Aligner(C.Structure):       # quicked_aligner_t, 72 bytes
    _fields_ = [("params", C.POINTER(Params)), ("mm_allocator", C.c_void_p), ("cigar", C.c_char_p), ("score", C.c_int),
                ("timer", C.c_void_p), ("timer_windowed_s", C.c_void_p), ("timer_windowed_l", C.c_void_p),
                ("timer_banded", C.c_void_p), ("timer_align", C.c_void_p)]


class Batch(C.Structure):         # qb200_batch_t
    _fields_ = [("seqs", C.c_void_p), ("seqs_bytes", C.c_int64), ("n_pairs", C.c_int64), ("pattern_off", C.c_void_p),
                ("pattern_len", C.c_void_p), ("text_off", C.c_void_p), ("text_len", C.c_void_p)]


class PackedBatch(C.Structure):   # qb200_packed_batch_t
    _fields_ = [("packed", C.c_void_p), ("n_chars", C.c_int64), ("n_pairs", C.c_int64), ("pattern_off", C.c_void_p),
                ("pattern_len", C.c_void_p), ("text_off", C.c_void_p), ("text_len", C.c_void_p), ("exc_pos", C.c_void_p),
                ("exc_chr", C.c_void_p), ("n_exc", C.c_int64)]


class Results(C.Structure):       # qb200_results_t
    _fields_ = [("score", C.c_void_p), ("status", C.c_void_p), ("cigar", C.c_void_p), ("cigar_capacity", C.c_int64),
                ("cigar_off", C.c_void_p), ("cigar_bytes", C.c_int64)]


class Stats(C.Structure):         # qb200_stats_t
    _fields_ = [(k, C.c_int64) for k in ("n_pairs", "kernel_launches", "word_steps", "word_steps_windowed",
                                         "word_steps_banded", "cells", "h2d_bytes", "d2h_bytes", "pairs_stage2",
                                         "pairs_stage3", "banded_tries", "hirschberg_splits", "leaves")] + \
               [(k, C.c_float) for k in ("ms_total", "ms_prepare", "ms_windowed_s", "ms_windowed_l", "ms_banded",
                                         "ms_align_fill", "ms_align_trace", "ms_cigar")] + [("matrix_bytes", C.c_int64)] + \
               [("ms_fused", C.c_float), ("leaves_punted", C.c_int32), ("pairs_fused", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load():
    """Load libquicked_b200.so (built in-tree by quicked_b200/build.py).  No fallback: raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m quicked_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback for the QuickEd GPU path.")
    L = C.CDLL(LIB_PATH)
    L.quicked_default_params.restype = Params
    L.quicked_new.argtypes = [C.POINTER(Aligner), C.POINTER(Params)]
    L.quicked_free.argtypes = [C.POINTER(Aligner)]
    L.quicked_align.argtypes = [C.POINTER(Aligner), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.quicked_status_msg.restype = C.c_char_p
    L.quicked_status_msg.argtypes = [C.c_int]
    L.quicked_check_error.restype = C.c_bool
    L.quicked_check_error.argtypes = [C.c_int]
    L.qb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.qb200_destroy.argtypes = [C.c_void_p]
    L.qb200_destroy.restype = None
    L.qb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.qb200_set_workspace_limit.argtypes = [C.c_void_p, C.c_size_t]
    L.qb200_last_error.restype = C.c_char_p
    L.qb200_last_error.argtypes = [C.c_void_p]
    for f in ("qb200_upload", "qb200_upload_device"):
        getattr(L, f).argtypes = [C.c_void_p, C.POINTER(Batch)]
    L.qb200_run.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.qb200_download.argtypes = [C.c_void_p, C.POINTER(Results)]
    L.qb200_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.qb200_get_bounds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    L.qb200_cigar_to_sam.restype = C.c_int64
    L.qb200_cigar_to_sam.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int64]
    L.qb200_align_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(Batch), C.POINTER(Results)]
    L.qb200_pack_batch.restype = C.c_int64
    L.qb200_pack_batch.argtypes = [C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
    L.qb200_upload_packed.argtypes = [C.c_void_p, C.POINTER(PackedBatch)]
    L.qb200_align_batch_packed.argtypes = [C.c_void_p, C.POINTER(Params), C.POINTER(PackedBatch), C.POINTER(Results)]
    L.qb200_generate_device.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32]
    L.qb200_batch_bytes.restype = C.c_int64
    L.qb200_batch_bytes.argtypes = [C.c_void_p]
    L.qb200_download_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.qb200_measure_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.qb200_host_alloc.restype = C.c_void_p
    L.qb200_host_alloc.argtypes = [C.c_size_t]
    L.qb200_host_free.argtypes = [C.c_void_p]
    L.qb200_host_free.restype = None
    L.qb200_generate_pairs.restype = C.c_int64
    L.qb200_generate_pairs.argtypes = [C.c_uint64, C.c_int64, C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]
    _lib = L
    return L


class QuickedException(Exception):
    def __init__(self, status):
        self.status = status
        super().__init__(load().quicked_status_msg(status).decode().strip())


def make_params(**kw):
    p = load().quicked_default_params()
    for k, v in kw.items():
        if k in ("hew_threshold", "hew_percentage"):
            if isinstance(v, int):
                v = (v, v)
            getattr(p, k)[0], getattr(p, k)[1] = v
        else:
            setattr(p, k, v)
    return p


class QuickedAligner:
    """Single-pair aligner with the reference binding's interface (bindings/cpp/quicked.hpp:46-73)."""

    def __init__(self):
        self._lib = load()
        self._params = self._lib.quicked_default_params()
        self._aligner = Aligner()
        st = self._lib.quicked_new(C.byref(self._aligner), C.byref(self._params))
        if self._lib.quicked_check_error(st):
            raise QuickedException(st)

    def __del__(self):
        try:
            self._lib.quicked_free(C.byref(self._aligner))
        except Exception:
            pass

    def align(self, pattern, text):
        pattern = pattern.encode() if isinstance(pattern, str) else bytes(pattern)
        text = text.encode() if isinstance(text, str) else bytes(text)
        st = self._lib.quicked_align(C.byref(self._aligner), pattern, len(pattern), text, len(text))
        self.status = st
        if self._lib.quicked_check_error(st):
            raise QuickedException(st)
        return st

    def setAlgorithm(self, algo): self._params.algo = int(algo)
    def setOnlyScore(self, v): self._params.only_score = bool(v)
    def setBandwidth(self, v): self._params.bandwidth = int(v)
    def setWindowSize(self, v): self._params.window_size = int(v)
    def setOverlapSize(self, v): self._params.overlap_size = int(v)
    def setForceScalar(self, v): self._params.force_scalar = bool(v)
    def setHEWThreshold(self, v): self._params.hew_threshold[0] = self._params.hew_threshold[1] = int(v)
    def setHEWPercentage(self, v): self._params.hew_percentage[0] = self._params.hew_percentage[1] = int(v)
    def getScore(self): return self._aligner.score
    def getCigar(self): return self._aligner.cigar.decode() if self._aligner.cigar else "NULL"


def pack_pairs(pairs):
    """[(pattern, text), ...] -> (seqs uint8[...], pattern_off, pattern_len, text_off, text_len) numpy arrays,
    the packed layout of qb200_batch_t (pattern i then text i, back to back)."""
    n = len(pairs)
    po = np.zeros(n, np.int64); to = np.zeros(n, np.int64)
    pl = np.zeros(n, np.int32); tl = np.zeros(n, np.int32)
    chunks, off = [], 0
    for i, (p, t) in enumerate(pairs):
        p = p.encode() if isinstance(p, str) else bytes(p)
        t = t.encode() if isinstance(t, str) else bytes(t)
        po[i], pl[i] = off, len(p); off += len(p)
        to[i], tl[i] = off, len(t); off += len(t)
        chunks.append(p); chunks.append(t)
    seqs = np.frombuffer(b"".join(chunks) + b"\0", dtype=np.uint8).copy()
    return seqs, po, pl, to, tl


class BatchAligner:
    """One GPU context (device buffers, stream).  align(pairs) -> (status[], score[], [cigar str])."""

    def __init__(self, device=0, stream=None, workspace_limit=None):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.qb200_create(C.byref(h), int(device))
        if rc == QB200_ERR_NO_DEVICE:
            raise RuntimeError("quicked_b200: no CUDA device — the GPU path has no CPU fallback")
        if rc != 0:
            raise RuntimeError(f"qb200_create failed rc={rc}")
        self._h = h
        if stream is not None:
            self._lib.qb200_set_stream(self._h, C.c_void_p(int(stream)))
        if workspace_limit:
            self._lib.qb200_set_workspace_limit(self._h, int(workspace_limit))
        self._keep = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed rc={rc}: {self._lib.qb200_last_error(self._h).decode()}")

    @staticmethod
    def _batch(seqs_ptr, seqs_bytes, n, po, pl, to, tl):
        return Batch(seqs_ptr, seqs_bytes, n, po, pl, to, tl)

    def upload_arrays(self, seqs, po, pl, to, tl):
        """numpy (host) arrays in the packed layout"""
        self._keep = (seqs, po, pl, to, tl)
        b = self._batch(seqs.ctypes.data, int(seqs.size), int(po.size), po.ctypes.data, pl.ctypes.data, to.ctypes.data,
                        tl.ctypes.data)
        self._n = int(po.size)
        self._check(self._lib.qb200_upload(self._h, C.byref(b)), "qb200_upload")

    def upload_device_ptrs(self, seqs_ptr, seqs_bytes, n, po_ptr, pl_ptr, to_ptr, tl_ptr):
        """device pointers (e.g. torch tensors' data_ptr()); the character buffer is used in place"""
        b = self._batch(seqs_ptr, seqs_bytes, n, po_ptr, pl_ptr, to_ptr, tl_ptr)
        self._n = int(n)
        self._check(self._lib.qb200_upload_device(self._h, C.byref(b)), "qb200_upload_device")

    def run(self, params=None, **kw):
        p = params if params is not None else make_params(**kw)
        self._check(self._lib.qb200_run(self._h, C.byref(p)), "qb200_run")

    def download(self, want_cigar=True):
        n = self._n
        score = np.empty(n, np.int32); status = np.empty(n, np.int32)
        off = np.zeros(n + 1, np.int64)
        r = Results(score.ctypes.data, status.ctypes.data, None, 0, off.ctypes.data, 0)
        rc = self._lib.qb200_download(self._h, C.byref(r))
        cig = None
        if rc == QB200_ERR_CAPACITY and want_cigar:
            cig = np.empty(int(r.cigar_bytes), np.uint8)
            r = Results(score.ctypes.data, status.ctypes.data, cig.ctypes.data, int(cig.size), off.ctypes.data, 0)
            rc = self._lib.qb200_download(self._h, C.byref(r))
        if rc not in (0, QB200_ERR_CAPACITY):
            self._check(rc, "qb200_download")
        return status, score, off, cig

    def int_peak_tops(self):
        v = C.c_double()
        self._check(self._lib.qb200_measure_int_peak(self._h, C.byref(v)), "qb200_measure_int_peak")
        return v.value

    def bounds(self):
        """stage-1 WindowEd(S) score and high-error-window count per pair of the last QUICKED run"""
        b = np.empty(self._n, np.int32); h = np.empty(self._n, np.int32)
        self._check(self._lib.qb200_get_bounds(self._h, b.ctypes.data, h.ctypes.data, self._n), "qb200_get_bounds")
        return b, h

    def stats(self):
        s = Stats()
        self._lib.qb200_get_stats(self._h, C.byref(s))
        return s.as_dict()

    def align_batch(self, pairs, params=None, **kw):
        """qb200_align_batch (host in / host out; pipelined for big jobs) -> list of (status, score, cigar-or-None)"""
        seqs, po, pl, to, tl = pack_pairs(pairs)
        n = int(po.size)
        p = params if params is not None else make_params(**kw)
        score = np.empty(n, np.int32); status = np.empty(n, np.int32); off = np.zeros(n + 1, np.int64)
        cig = np.zeros(int(seqs.size) // 2 + 1024, np.uint8)
        b = self._batch(seqs.ctypes.data, int(seqs.size), n, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
        r = Results(score.ctypes.data, status.ctypes.data, cig.ctypes.data, int(cig.size), off.ctypes.data, 0)
        rc = self._lib.qb200_align_batch(self._h, C.byref(p), C.byref(b), C.byref(r))
        if rc == QB200_ERR_CAPACITY:
            cig = np.zeros(int(r.cigar_bytes) + 16, np.uint8)
            r = Results(score.ctypes.data, status.ctypes.data, cig.ctypes.data, int(cig.size), off.ctypes.data, 0)
            rc = self._lib.qb200_align_batch(self._h, C.byref(p), C.byref(b), C.byref(r))
        self._check(rc, "qb200_align_batch")
        raw = cig.tobytes()
        out = []
        for i in range(n):
            c = raw[off[i]:off[i + 1] - 1].decode() if off[i + 1] - off[i] > 1 else None
            out.append((int(status[i]), int(score[i]), c))
        return out

    def generate_device(self, seed, n_pairs, length, error, first=0, indels=None):
        """qb200_generate_device: the batch is generated by a kernel in this context's device buffers"""
        ind = indels or (0, 0)
        self._n = int(n_pairs)
        self._check(self._lib.qb200_generate_device(self._h, int(seed), int(first), int(n_pairs), int(length), float(error), int(ind[0]), int(ind[1])),
                    "qb200_generate_device")

    def download_batch(self):
        """the batch held by the context -> (seqs, po, pl, to, tl) numpy arrays"""
        n = self._n
        seqs = np.zeros(int(self._lib.qb200_batch_bytes(self._h)), np.uint8)
        po = np.zeros(n, np.int64); to = np.zeros(n, np.int64); pl = np.zeros(n, np.int32); tl = np.zeros(n, np.int32)
        self._check(self._lib.qb200_download_batch(self._h, seqs.ctypes.data, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data),
                    "qb200_download_batch")
        return seqs, po, pl, to, tl

    def upload_packed_arrays(self, packed, n_chars, po, pl, to, tl, exc_pos, exc_chr):
        """2-bit packed stream + exception list (pack_2bit) -> qb200_upload_packed"""
        self._keep = (packed, po, pl, to, tl, exc_pos, exc_chr)
        b = PackedBatch(packed.ctypes.data, int(n_chars), int(po.size), po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data,
                        exc_pos.ctypes.data if exc_pos.size else None, exc_chr.ctypes.data if exc_chr.size else None, int(exc_pos.size))
        self._n = int(po.size)
        self._check(self._lib.qb200_upload_packed(self._h, C.byref(b)), "qb200_upload_packed")

    def align_packed(self, pairs, params=None, **kw):
        """align(pairs) through the 2-bit packed upload"""
        seqs, po, pl, to, tl = pack_pairs(pairs)
        packed, ep, ec = pack_2bit(seqs, po, pl, to, tl)
        self.upload_packed_arrays(packed, int(seqs.size), po, pl, to, tl, ep, ec)
        return self._finish(len(pairs), params, **kw)

    def align_batch_packed(self, pairs, params=None, **kw):
        """qb200_align_batch_packed (2-bit packed host input, host out; pipelined for big jobs)"""
        seqs, po, pl, to, tl = pack_pairs(pairs)
        packed, ep, ec = pack_2bit(seqs, po, pl, to, tl)
        n = int(po.size)
        p = params if params is not None else make_params(**kw)
        score = np.empty(n, np.int32); status = np.empty(n, np.int32); off = np.zeros(n + 1, np.int64)
        b = PackedBatch(packed.ctypes.data, int(seqs.size), n, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data,
                        ep.ctypes.data if ep.size else None, ec.ctypes.data if ec.size else None, int(ep.size))
        cig = np.zeros(int(seqs.size) // 2 + 1024, np.uint8)
        for _ in range(2):
            r = Results(score.ctypes.data, status.ctypes.data, cig.ctypes.data, int(cig.size), off.ctypes.data, 0)
            rc = self._lib.qb200_align_batch_packed(self._h, C.byref(p), C.byref(b), C.byref(r))
            if rc != QB200_ERR_CAPACITY:
                break
            cig = np.zeros(int(r.cigar_bytes) + 16, np.uint8)
        self._check(rc, "qb200_align_batch_packed")
        raw = cig.tobytes()
        return [(int(status[i]), int(score[i]), raw[off[i]:off[i + 1] - 1].decode() if off[i + 1] - off[i] > 1 else None) for i in range(n)]

    def align(self, pairs, params=None, **kw):
        """-> list of (status, score, cigar-or-None), one tuple per pair"""
        self.upload_arrays(*pack_pairs(pairs))
        return self._finish(len(pairs), params, **kw)

    def _finish(self, n_pairs, params=None, **kw):
        self.run(params, **kw)
        status, score, off, cig = self.download()
        out = []
        raw = cig.tobytes() if cig is not None else b""
        for i in range(n_pairs):
            c = None
            if cig is not None and off[i + 1] - off[i] > 1:
                c = raw[off[i]:off[i + 1] - 1].decode()
            out.append((int(status[i]), int(score[i]), c))
        return out


def pack_2bit(seqs, po, pl, to, tl, threads=0):
    """Host packer qb200_pack_batch: the ASCII batch (numpy arrays of pack_pairs / generate_pairs_native) ->
    (packed uint8[(n+3)//4 + 8], exc_pos int64[], exc_chr uint8[]); the offset / length arrays serve both formats."""
    L = load()
    b = Batch(seqs.ctypes.data, int(seqs.size), int(po.size), po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    packed = np.zeros((int(seqs.size) + 3) // 4 + 8, np.uint8)
    cap = 1024
    while True:
        ep = np.zeros(cap, np.int64); ec = np.zeros(cap, np.uint8)
        ne = L.qb200_pack_batch(C.byref(b), packed.ctypes.data, ep.ctypes.data, ec.ctypes.data, cap, int(threads))
        if ne >= 0:
            return packed, ep[:ne].copy(), ec[:ne].copy()
        if ne <= -100:
            raise RuntimeError(f"qb200_pack_batch rc={ne}")
        cap = -ne + 16


def unpack_2bit(packed, n_chars, exc_pos, exc_chr):
    """numpy restatement of the device unpack (k_unpack2 + k_patch_exceptions): the characters the kernels will see"""
    idx = np.arange(n_chars)
    codes = (packed[idx >> 2] >> (2 * (idx & 3))) & 3
    out = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
    out[exc_pos] = exc_chr
    return out


def cigar_to_sam(cigar, show_mismatches=False):
    """SAM-style CIGAR of one alignment (reference cigar_sprint_SAM_CIGAR)"""
    lib = load()
    c = cigar.encode() if isinstance(cigar, str) else cigar
    buf = C.create_string_buffer(len(c) + 16)
    n = lib.qb200_cigar_to_sam(c, int(bool(show_mismatches)), buf, len(buf))
    if n < 0:
        buf = C.create_string_buffer(-n)
        n = lib.qb200_cigar_to_sam(c, int(bool(show_mismatches)), buf, len(buf))
    return buf.raw[:n].decode()


def generate_pairs_native(seed, n_pairs, length, error, first=0, indels=None):
    """Seeded generate_dataset twin in C (qb200_generate_pairs_ex): pairs [first, first + n_pairs) of job `seed`,
    optional indels=(num, length) like the reference's --indels.  -> (seqs, po, pl, to, tl) numpy arrays."""
    import math
    L = load()
    nerr = int(error) if error >= 1.0 else int(math.ceil(np.float32(length) * np.float32(error)))
    stride = 2 * length + nerr + 2
    total = (n_pairs * stride + 15) // 16 * 16
    seqs = np.zeros(total, np.uint8)
    po = np.zeros(n_pairs, np.int64); to = np.zeros(n_pairs, np.int64)
    pl = np.zeros(n_pairs, np.int32); tl = np.zeros(n_pairs, np.int32)
    L.qb200_generate_pairs_ex.restype = C.c_int64
    L.qb200_generate_pairs_ex.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    ind = indels or (0, 0)
    rc = L.qb200_generate_pairs_ex(seed, first, n_pairs, length, float(error), int(ind[0]), int(ind[1]), seqs.ctypes.data,
                                   po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    if rc < 0:
        raise RuntimeError(f"qb200_generate_pairs_ex rc={rc}")
    return seqs, po, pl, to, tl
