"""oracle/harness.py — ctypes access to the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * `Oracle`    -> oracle/libqoracle.so   (our CPU restatement, oracle/quicked_oracle.c)
  * `Reference` -> oracle/_ref/libquicked_ref.so (the unmodified reference, built by oracle/Makefile from
                   /root/reference; the built .so travels to the GPU box, the sources do not)

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this module.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libqoracle.so")
REF_SO = os.path.join(HERE, "_ref", "libquicked_ref.so")
REF_BATCH_SO = os.path.join(HERE, "_ref", "libref_batch.so")

QUICKED, WINDOWED, BANDED, HIRSCHBERG = 0, 1, 2, 3


def build(ref=True):
    """Compile the checkers (idempotent).  `make ref` is a no-op when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        # the reference's unmodified callers on the product library (drop-in evidence, tests/test_reference_callers.py)
        subprocess.run(["make", "-s", "-C", HERE, "refcallers"], check=True)


class QoParams(C.Structure):
    _fields_ = [("algo", C.c_int), ("bandwidth", C.c_uint), ("window_size", C.c_uint), ("overlap_size", C.c_uint),
                ("hew_threshold", C.c_uint * 2), ("hew_percentage", C.c_uint * 2), ("only_score", C.c_int),
                ("force_scalar", C.c_int)]


class QoResult(C.Structure):
    _fields_ = [("status", C.c_int), ("score", C.c_int64), ("cigar", C.c_void_p), ("bound_ws", C.c_int64),
                ("bound_final", C.c_int64), ("stage", C.c_int), ("banded_tries", C.c_int), ("splits", C.c_int), ("ref_undefined", C.c_int),
                ("word_steps", C.c_uint64), ("word_steps_windowed", C.c_uint64), ("word_steps_banded", C.c_uint64)]


class BandGeom(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("k", "d", "rel", "prolog", "B_cigar", "B_score", "fin")]


def _b(s):
    return s if isinstance(s, (bytes, bytearray)) else s.encode()


class Oracle:
    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(ORACLE_SO)
        L.qo_default_params.restype = QoParams
        L.qo_align.argtypes = [C.POINTER(QoParams), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(QoResult)]
        L.qo_free_result.argtypes = [C.POINTER(QoResult)]
        L.qo_status_msg.restype = C.c_char_p
        L.qo_band_geometry.restype = BandGeom
        L.qo_band_geometry.argtypes = [C.c_int64] * 3
        L.qo_banded_score.restype = C.c_int64
        L.qo_banded_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.qo_windowed_score.restype = C.c_int64
        L.qo_windowed_score.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_int64)]
        L.qo_align_ops.restype = C.c_void_p
        L.qo_align_ops.argtypes = [C.POINTER(QoParams), C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int),
                                   C.POINTER(C.c_int64)]
        self.libc = C.CDLL(None)
        self.libc.free.argtypes = [C.c_void_p]

    def params(self, **kw):
        p = self.lib.qo_default_params()
        for k, v in kw.items():
            if k in ("hew_threshold", "hew_percentage"):
                getattr(p, k)[0], getattr(p, k)[1] = v
            else:
                setattr(p, k, int(v))
        return p

    def align(self, pattern, text, full=False, **kw):
        """-> (status, score, cigar or None) [+ diagnostics dict when full]"""
        p = self.params(**kw)
        pattern, text = _b(pattern), _b(text)
        r = QoResult()
        self.lib.qo_align(C.byref(p), pattern, len(pattern), text, len(text), C.byref(r))
        cig = C.string_at(r.cigar).decode() if r.cigar else None
        out = (r.status, r.score, cig)
        if full:
            out = out + ({k: getattr(r, k) for k in ("bound_ws", "bound_final", "stage", "banded_tries", "splits", "ref_undefined",
                                                     "word_steps", "word_steps_windowed", "word_steps_banded")},)
        self.lib.qo_free_result(C.byref(r))
        return out

    def align_ops(self, pattern, text, **kw):
        p = self.params(**kw)
        pattern, text = _b(pattern), _b(text)
        st, sc = C.c_int(), C.c_int64()
        ptr = self.lib.qo_align_ops(C.byref(p), pattern, len(pattern), text, len(text), C.byref(st), C.byref(sc))
        ops = C.string_at(ptr).decode() if ptr else None
        if ptr:
            self.libc.free(ptr)
        return st.value, sc.value, ops

    def status_msg(self, st):
        return self.lib.qo_status_msg(C.c_int(st)).decode()

    def band_geometry(self, m, n, cutoff):
        return self.lib.qo_band_geometry(m, n, cutoff)

    def banded_score(self, pattern, text, cutoff, finish=None):
        pattern, text = _b(pattern), _b(text)
        lo, hi = C.c_int64(), C.c_int64()
        s = self.lib.qo_banded_score(pattern, len(pattern), text, len(text), cutoff,
                                     len(text) if finish is None else finish, None, None, None, C.byref(lo), C.byref(hi))
        return s, lo.value, hi.value

    def windowed_score(self, pattern, text, W, O, hew_threshold, sse):
        pattern, text = _b(pattern), _b(text)
        hew = C.c_int64()
        s = self.lib.qo_windowed_score(pattern, len(pattern), text, len(text), W, O, hew_threshold, int(sse), C.byref(hew))
        return s, hew.value


# ---- the unmodified reference, through its own public API (quicked/quicked.h) ----
class RefParams(C.Structure):   # quicked/quicked.h:43-54; x86-64 layout: 48 bytes
    _fields_ = [("algo", C.c_int), ("bandwidth", C.c_uint), ("window_size", C.c_uint), ("overlap_size", C.c_uint),
                ("hew_threshold", C.c_uint * 2), ("hew_percentage", C.c_uint * 2), ("only_score", C.c_bool),
                ("force_scalar", C.c_bool), ("external_timer", C.c_bool), ("external_allocator", C.c_void_p)]


class RefAligner(C.Structure):  # quicked/quicked.h:56-67; 72 bytes
    _fields_ = [("params", C.POINTER(RefParams)), ("mm_allocator", C.c_void_p), ("cigar", C.c_char_p),
                ("score", C.c_int), ("timer", C.c_void_p), ("timer_windowed_s", C.c_void_p),
                ("timer_windowed_l", C.c_void_p), ("timer_banded", C.c_void_p), ("timer_align", C.c_void_p)]


class Reference:
    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = self.lib = C.CDLL(path)
        L.quicked_default_params.restype = RefParams
        L.quicked_new.argtypes = [C.POINTER(RefAligner), C.POINTER(RefParams)]
        L.quicked_align.argtypes = [C.POINTER(RefAligner), C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        L.quicked_free.argtypes = [C.POINTER(RefAligner)]
        L.quicked_status_msg.restype = C.c_char_p
        L.quicked_status_msg.argtypes = [C.c_int]
        L.quicked_check_error.restype = C.c_bool
        L.quicked_check_error.argtypes = [C.c_int]

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def params(self, **kw):
        p = self.lib.quicked_default_params()
        for k, v in kw.items():
            if k in ("hew_threshold", "hew_percentage"):
                getattr(p, k)[0], getattr(p, k)[1] = v
            else:
                setattr(p, k, v)
        return p

    def align(self, pattern, text, **kw):
        """one quicked_new / quicked_align / quicked_free round trip -> (status, score, cigar or None)"""
        p = self.params(**kw)
        a = RefAligner()
        self.lib.quicked_new(C.byref(a), C.byref(p))
        pattern, text = _b(pattern), _b(text)
        # NUL-terminated copies: the SSE window reads text[n] (bpm_windowed.c:361); every reference caller
        # provides a terminator there (align_benchmark.c:95-97, argv strings), and so do we.
        st = self.lib.quicked_align(C.byref(a), pattern, len(pattern), text, len(text))
        score, cig = a.score, (a.cigar.decode() if a.cigar else None)
        self.lib.quicked_free(C.byref(a))
        return st, score, cig

    def status_msg(self, st):
        return self.lib.quicked_status_msg(st).decode()


def cpu_batch_align(seqs, po, pl, to, tl, threads, algo=0, bandwidth=15, window_size=9, overlap_size=1, only_score=False,
                    force_scalar=False, want_scores=False):
    """Time-critical CPU baseline: ONE native call aligns the whole packed batch on `threads` pthreads, through the
    unmodified reference (oracle/_ref/libref_batch.so, the per-thread loop of the reference's align_benchmark) when it is
    built, else through the oracle port.  numpy arrays in the layout of qb200_batch_t.  -> (kind, cigar_bytes, scores|None)"""
    import numpy as np
    n = int(po.size)
    scores = np.zeros(n, np.int32) if want_scores else None
    sp = scores.ctypes.data if want_scores else None
    if os.path.exists(REF_BATCH_SO):
        L = C.CDLL(REF_BATCH_SO)
        L.ref_batch_align.restype = C.c_int64
        L.ref_batch_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint,
                                      C.c_uint, C.c_uint, C.c_int, C.c_int, C.c_void_p]
        b = L.ref_batch_align(seqs.ctypes.data, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data, n, int(threads), int(algo),
                              int(bandwidth), int(window_size), int(overlap_size), int(only_score), int(force_scalar), sp)
        return "reference", int(b), scores
    o = Oracle()
    o.lib.qo_batch_align.restype = C.c_int64
    o.lib.qo_batch_align.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.POINTER(QoParams), C.c_void_p]
    p = o.params(algo=algo, bandwidth=bandwidth, window_size=window_size, overlap_size=overlap_size, only_score=only_score, force_scalar=force_scalar)
    b = o.lib.qo_batch_align(seqs.ctypes.data, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data, n, int(threads), C.byref(p), sp)
    return "port", int(b), scores


def generate_pairs(seed, n_pairs, length, error, first=0, indels=None, threads=None):
    """Checker-side twin of quicked_b200.generate_pairs_native (oracle/datagen.c: same model, same random streams):
    pairs [first, first + n_pairs) of job `seed` in the packed layout.  Lets bench.py --impl reference make its data
    without loading the product library.  -> (seqs, po, pl, to, tl) numpy arrays."""
    import math
    import numpy as np
    if not os.path.exists(ORACLE_SO):
        build(ref=False)
    L = C.CDLL(ORACLE_SO)
    L.qo_generate_pairs.restype = C.c_int64
    L.qo_generate_pairs.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    nerr = int(error) if error >= 1.0 else int(math.ceil(np.float32(length) * np.float32(error)))
    stride = 2 * length + nerr + 2
    seqs = np.zeros((n_pairs * stride + 15) // 16 * 16, np.uint8)
    po = np.zeros(n_pairs, np.int64); to = np.zeros(n_pairs, np.int64)
    pl = np.zeros(n_pairs, np.int32); tl = np.zeros(n_pairs, np.int32)
    ind = indels or (0, 0)
    rc = L.qo_generate_pairs(seed, first, n_pairs, length, float(error), int(ind[0]), int(ind[1]), int(threads or os.cpu_count() or 1),
                             seqs.ctypes.data, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    if rc < 0:
        raise RuntimeError("qo_generate_pairs failed")
    return seqs, po, pl, to, tl
