"""TEST INFRASTRUCTURE ONLY -- ctypes bindings of the two CPU checkers.

* ``Ref``  : oracle/_ref/libmbpol_ref.so, the reference's own hot-path sources compiled unmodified
             (oracle/Makefile, wrappers oracle/ref_wrapper_*.cpp).  The parity authority.
* ``Port`` : oracle/liboracle.so, our plain-C restatement (oracle/mbpol_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product package (mbpol_openmm_plugin_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libmbpol_ref.so")
PORT_SO = os.path.join(HERE, "liboracle.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def canonical(tuples):
    """Sort rows lexicographically (ascending first, second[, third]) -- SURVEY.md App. B canonical order."""
    t = np.asarray(tuples, dtype=np.int64)
    if t.size == 0:
        return t.astype(np.int32)
    t = t.reshape(len(t), -1)
    order = np.lexsort(tuple(t[:, k] for k in range(t.shape[1] - 1, -1, -1)))
    return t[order].astype(np.int32)


class Ref:
    """The reference's own code (unmodified), behind a C ABI."""

    kind = "reference"

    def __init__(self, path=REF_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built; run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(path)
        L.ref_triplets.restype = C.c_long
        L.ref_triplets.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, _up, C.c_long]
        L.ref_pairs.restype = C.c_long
        L.ref_pairs.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, _up, C.c_long]
        for fn in (L.ref_eval_2b, L.ref_eval_3b):
            fn.restype = C.c_double
            fn.argtypes = [_dp, _dp, C.c_long, _ip, C.c_int, _up, C.c_long, _dp, C.c_int, C.c_double]
        for fn in (L.ref_each_2b, L.ref_each_3b):
            fn.restype = None
            fn.argtypes = [_dp, C.c_int, _up, C.c_long, _dp, C.c_int, _dp, _dp]
        for fn in (L.ref_poly2b, L.ref_poly3b):
            fn.restype = C.c_double
            fn.argtypes = [_dp, _dp]
        if hasattr(L, "ref_eval_1b"):
            L.ref_eval_1b.restype = C.c_double
            L.ref_eval_1b.argtypes = [_dp, _dp, C.c_long, _ip, C.c_int, _dp, C.c_int]
        L.ref_ncoef2b.restype = L.ref_ncoef3b.restype = C.c_int
        L.ref_coef2b.restype = L.ref_coef3b.restype = _dp

    # -- lists -------------------------------------------------------------------------------
    def _list(self, fn, width, posO, box, pbc, cutoff):
        posO = _f64(posO)
        box = _f64(box if box is not None else [0, 0, 0])
        n = len(posO)
        cap = max(1024, 64 * n)
        while True:
            out = np.empty((cap, width), dtype=np.uint32)
            m = fn(_d(posO), n, _d(box), int(bool(pbc)), float(cutoff), out.ctypes.data_as(_up), cap)
            if m <= cap:
                return out[:m].astype(np.int32)
            cap = m

    def pairs(self, posO, box, pbc, cutoff):
        """(first=i, second=j<i) pairs with min-image d2 <= cutoff2, reference emission order."""
        return self._list(self.lib.ref_pairs, 2, posO, box, pbc, cutoff)

    def triplets(self, posO, box, pbc, cutoff):
        return self._list(self.lib.ref_triplets, 3, posO, box, pbc, cutoff)

    # -- forces ------------------------------------------------------------------------------
    def _eval(self, fn, pos, idx, tuples, box, pbc, cutoff):
        pos = _f64(pos)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        tup = np.ascontiguousarray(tuples, dtype=np.uint32)
        box = _f64(box if box is not None else [0, 0, 0])
        f = np.zeros_like(pos)
        e = fn(_d(pos), _d(f), len(pos), idx.ctypes.data_as(_ip), len(idx), tup.ctypes.data_as(_up), len(tup),
               _d(box), int(bool(pbc)), float(cutoff))
        return e, f

    def eval_2b(self, pos, idx, pairs, box, pbc, cutoff=1e10):
        """E (kJ/mol), F (kJ/mol/nm, n_particles x 3) over the given pair list."""
        return self._eval(self.lib.ref_eval_2b, pos, idx, pairs, box, pbc, cutoff)

    def eval_3b(self, pos, idx, trip, box, pbc, cutoff=1e10):
        return self._eval(self.lib.ref_eval_3b, pos, idx, trip, box, pbc, cutoff)

    def eval_1b(self, pos, idx, box, pbc):
        """one-body term: E (kJ/mol), F (kJ/mol/nm) summed over the molecules in idx"""
        pos = _f64(pos)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        box = _f64(box if box is not None else [0, 0, 0])
        f = np.zeros_like(pos)
        e = self.lib.ref_eval_1b(_d(pos), _d(f), len(pos), idx.ctypes.data_as(_ip), len(idx), _d(box), int(bool(pbc)))
        return e, f

    def _each(self, fn, width, pos_mol, tuples, box, pbc):
        pos = _f64(pos_mol).reshape(-1, 3)
        tup = np.ascontiguousarray(tuples, dtype=np.uint32)
        box = _f64(box if box is not None else [0, 0, 0])
        e = np.zeros(len(tup))
        f = np.zeros((len(tup), 3 * width, 3))
        fn(_d(pos), len(pos) // 3, tup.ctypes.data_as(_up), len(tup), _d(box), int(bool(pbc)), _d(e), _d(f))
        return e, f

    def each_2b(self, pos_mol, pairs, box, pbc):
        """per-pair E and 6x3 forces; pos_mol = (n_mol,3,3) O,H1,H2 in nm."""
        return self._each(self.lib.ref_each_2b, 2, pos_mol, pairs, box, pbc)

    def each_3b(self, pos_mol, trip, box, pbc):
        return self._each(self.lib.ref_each_3b, 3, pos_mol, trip, box, pbc)

    # -- polynomials -------------------------------------------------------------------------
    def poly2b(self, x):
        x = _f64(x)
        g = np.zeros(31)
        return self.lib.ref_poly2b(_d(x), _d(g)), g

    def poly3b(self, x):
        x = _f64(x)
        g = np.zeros(36)
        return self.lib.ref_poly3b(_d(x), _d(g)), g

    def coef2b(self):
        return np.ctypeslib.as_array(self.lib.ref_coef2b(), shape=(self.lib.ref_ncoef2b(),)).copy()

    def coef3b(self):
        return np.ctypeslib.as_array(self.lib.ref_coef3b(), shape=(self.lib.ref_ncoef3b(),)).copy()


def have_ref():
    return os.path.exists(REF_SO)


class Port:
    """Our plain-C restatement (oracle/mbpol_oracle.c -> oracle/liboracle.so).  kind = "port"."""

    kind = "port"

    def __init__(self, path=PORT_SO):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not built; run `make -C oracle port`")
        L = self.lib = C.CDLL(path)
        data = os.path.join(HERE, "..", "mbpol_openmm_plugin_b200", "data")
        self._keep = []
        u16p, u8p = C.POINTER(C.c_uint16), C.POINTER(C.c_uint8)
        L.port_set_tables.argtypes = [C.c_int, C.c_int, C.c_int, _dp, u16p, u8p]
        L.port_set_tables.restype = None
        for order, name in ((2, "pip2b"), (3, "pip3b")):
            t = np.load(os.path.join(data, name + ".npz"))
            coef = np.ascontiguousarray(t["coef"], dtype=np.float64)
            mc = np.ascontiguousarray(t["mono_coef"], dtype=np.uint16)
            mv = np.ascontiguousarray(t["mono_vars"], dtype=np.uint8)
            self._keep += [coef, mc, mv]
            L.port_set_tables(order, int(t["n_vars"]), len(mc), _d(coef), mc.ctypes.data_as(u16p), mv.ctypes.data_as(u8p))
        ob = np.load(os.path.join(data, "onebody.npz"))
        sc = np.array([float(ob[k]) for k in ("reoh", "b1", "roh", "alphaoh", "deohA", "phh1A", "phh2", "f5z",
                                              "cm1_kcalmol")])
        tabs = [np.ascontiguousarray(ob["c5z"], dtype=np.float64)] + \
            [np.ascontiguousarray(ob[k], dtype=np.int32) for k in ("idx1", "idx2", "idx3")]
        L.port_set_onebody.argtypes = [_dp, _ip, _ip, _ip, _dp]
        L.port_set_onebody.restype = None
        L.port_set_onebody(_d(tabs[0]), *[t.ctypes.data_as(_ip) for t in tabs[1:]], _d(sc))
        L.port_eval_1b.restype = C.c_double
        L.port_eval_1b.argtypes = [_dp, _dp, _ip, C.c_int, _dp, C.c_int]
        L.port_poly.restype = C.c_double
        L.port_poly.argtypes = [C.c_int, _dp, _dp]
        for fn in (L.port_pairs, L.port_triplets):
            fn.restype = C.c_long
            fn.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_double, _ip, C.c_long]
        for fn in (L.port_eval_2b, L.port_eval_3b):
            fn.restype = C.c_double
            fn.argtypes = [_dp, _dp, _ip, _ip, C.c_long, _dp, C.c_int, _dp, _dp]

    def poly(self, order, x):
        x = _f64(x)
        g = np.zeros(len(x))
        return self.lib.port_poly(order, _d(x), _d(g)), g

    def _list(self, fn, width, posO, box, pbc, cutoff):
        posO = _f64(posO)
        box = _f64(box if box is not None else [0, 0, 0])
        n = len(posO)
        cap = max(1024, 64 * n)
        while True:
            out = np.empty((cap, width), dtype=np.int32)
            m = fn(_d(posO), n, _d(box), int(bool(pbc)), float(cutoff), out.ctypes.data_as(_ip), cap)
            if m <= cap:
                return out[:m]
            cap = m

    def pairs(self, posO, box, pbc, cutoff):
        return self._list(self.lib.port_pairs, 2, posO, box, pbc, cutoff)

    def triplets(self, posO, box, pbc, cutoff):
        return self._list(self.lib.port_triplets, 3, posO, box, pbc, cutoff)

    def _eval(self, fn, width, pos, idx, tuples, box, pbc, each):
        pos = _f64(pos)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        tup = np.ascontiguousarray(tuples, dtype=np.int32)
        box = _f64(box if box is not None else [0, 0, 0])
        f = np.zeros_like(pos)
        e_each = np.zeros(len(tup)) if each else None
        f_each = np.zeros((len(tup), 3 * width, 3)) if each else None
        e = fn(_d(pos), _d(f), idx.ctypes.data_as(_ip), tup.ctypes.data_as(_ip), len(tup), _d(box), int(bool(pbc)),
               _d(e_each) if each else None, _d(f_each) if each else None)
        return (e, f, e_each, f_each) if each else (e, f)

    def eval_2b(self, pos, idx, pairs, box, pbc, cutoff=None, each=False):
        return self._eval(self.lib.port_eval_2b, 2, pos, idx, pairs, box, pbc, each)

    def eval_3b(self, pos, idx, trip, box, pbc, cutoff=None, each=False):
        return self._eval(self.lib.port_eval_3b, 3, pos, idx, trip, box, pbc, each)

    def eval_1b(self, pos, idx, box, pbc):
        pos = _f64(pos)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        box = _f64(box if box is not None else [0, 0, 0])
        f = np.zeros_like(pos)
        e = self.lib.port_eval_1b(_d(pos), _d(f), idx.ctypes.data_as(_ip), len(idx), _d(box), int(bool(pbc)))
        return e, f


def have_port():
    return os.path.exists(PORT_SO)


def eval_parallel(oracle, order, pos, idx, tuples, box, pbc, threads=None, chunk=20000):
    """E and F of a long tuple list, the checker fanned out over host threads (ctypes releases the GIL; the reference
    code keeps no state between calls).  Chunks are summed in list order, so the result does not depend on `threads`.
    Used for the whole-box comparison at 262,144 waters (about 330 core-seconds) and by bench.py's reference arm."""
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or len(os.sched_getaffinity(0))
    fn = oracle.eval_2b if order == 2 else oracle.eval_3b
    tuples = np.ascontiguousarray(tuples)
    parts = [tuples[i:i + chunk] for i in range(0, len(tuples), chunk)]
    e_tot, f_tot = 0.0, np.zeros_like(_f64(pos))
    with ThreadPoolExecutor(max_workers=threads) as ex:
        for res in ex.map(lambda t: fn(pos, idx, t, box, pbc)[:2], parts):
            e_tot += res[0]
            f_tot += res[1]
    return e_tot, f_tot


def best():
    """The strongest checker available: the reference's own code if it was built, else the port."""
    return Ref() if have_ref() else Port()
