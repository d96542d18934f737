"""ctypes binding of oracle/libhole_ref.so (TEST INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libhole_ref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        _lib = C.CDLL(_LIB)
        _lib.hole_ref_train_step.restype = C.c_double
        _lib.hole_ref_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def threads():
    return int(load().hole_ref_threads())


def set_threads(n):
    """OpenMP threads of the port (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib = load()
    lib.hole_ref_set_threads.argtypes = [C.c_int]
    lib.hole_ref_set_threads.restype = None
    lib.hole_ref_set_threads(int(n))


def score(E, triples):
    E = np.ascontiguousarray(E, np.float32)
    tr = np.ascontiguousarray(triples, np.int32)
    out = np.empty(len(tr), np.float32)
    load().hole_ref_score(_p(E), C.c_int(E.shape[1]), _p(tr), C.c_int64(len(tr)), _p(out))
    return out


class TrainScratch:
    def __init__(self, B, D):
        self.G = np.empty((6 * B, D), np.float32)
        self.idx = np.empty(6 * B, np.int32)
        self.loss = np.empty(B, np.float32)


def train_step(E, pos, neg_ent, side, margin, lr, scratch=None):
    """In place on E (float32 C-contiguous).  Returns (loss_sum, loss[B])."""
    assert E.dtype == np.float32 and E.flags.c_contiguous
    pos = np.ascontiguousarray(pos, np.int32)
    neg = np.ascontiguousarray(neg_ent, np.int32)
    B, D = len(pos), E.shape[1]
    sc = scratch or TrainScratch(B, D)
    tot = load().hole_ref_train_step(_p(E), C.c_int64(E.shape[0]), C.c_int(D), _p(pos), _p(neg),
                                     C.c_int(int(side)), C.c_int64(B), C.c_float(margin),
                                     C.c_float(lr), _p(sc.loss), _p(sc.G), _p(sc.idx))
    return float(tot), sc.loss


def logloss_step(E, pos, neg_ents, sides, lr, l2=0.0):
    """C port of hole_oracle.logloss_step, in place on E (float32 C-contiguous).
    Returns (loss [(1+k), B], l2_loss)."""
    assert E.dtype == np.float32 and E.flags.c_contiguous
    pos = np.ascontiguousarray(pos, np.int32)
    neg = np.ascontiguousarray(np.stack([np.asarray(n) for n in neg_ents]), np.int32)
    sd = np.ascontiguousarray(sides, np.int32)
    k, B, D = len(sd), len(pos), E.shape[1]
    loss = np.empty(((1 + k), B), np.float32)
    G = np.empty(((1 + k) * 3 * B, D), np.float32)
    idx = np.empty((1 + k) * 3 * B, np.int32)
    l2_loss = C.c_double(0.0)
    lib = load()
    lib.hole_ref_logloss_step.restype = None
    lib.hole_ref_logloss_step(_p(E), C.c_int64(E.shape[0]), C.c_int(D), _p(pos), _p(neg), _p(sd), C.c_int(k),
                              C.c_int64(B), C.c_float(lr), C.c_float(l2), _p(loss), C.byref(l2_loss), _p(G),
                              _p(idx))
    return loss, l2_loss.value


def clip_rows(E):
    E = np.ascontiguousarray(E, np.float32)
    Y = np.empty_like(E)
    load().hole_ref_clip_rows(_p(E), C.c_int64(E.shape[0]), C.c_int(E.shape[1]), _p(Y))
    return Y


def rank(Yc, cand_begin, qv, true_id, filter_off=None, filter_ids=None):
    Yc = np.ascontiguousarray(Yc, np.float32)
    qv = np.ascontiguousarray(qv, np.float32)
    tid = np.ascontiguousarray(true_id, np.int32)
    Q = len(qv)
    raw = np.empty(Q, np.int32)
    filt = np.empty(Q, np.int32)
    fo = fi = None
    if filter_off is not None:
        fo = np.ascontiguousarray(filter_off, np.int64)
        fi = np.ascontiguousarray(filter_ids, np.int32)
    load().hole_ref_rank(_p(Yc), C.c_int64(cand_begin), C.c_int64(cand_begin + len(Yc)),
                         C.c_int(Yc.shape[1]), _p(qv), _p(tid), C.c_int64(Q),
                         _p(fo) if fo is not None else None, _p(fi) if fi is not None else None,
                         _p(raw), _p(filt))
    return raw, filt
