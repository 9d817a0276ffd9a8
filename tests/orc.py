"""ctypes binding of the CPU oracle (oracle/dcp_oracle.c) -- test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_ODIR = os.path.join(_ROOT, "oracle")

ST_R, ST_S, ST_N, ST_B, ST_E, ST_J, ST_C, ST_T = [0xC000 | i for i in range(8)]
ENTRY_UNIFORM, ENTRY_OCCUPANCY = 1, 2
NTAB = 1364


def build():
    subprocess.check_call(["make", "-s", "-C", _ODIR], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _load(name):
    path = os.path.join(_ODIR, name)
    if not os.path.exists(path):
        build()
    try:
        return C.CDLL(path)
    except OSError:
        build()
        return C.CDLL(path)


class Oracle:
    """One flavour (float32 or float64 DP arithmetic) of the oracle."""

    def __init__(self, double=False):
        self.lib = _load("liborc_f64.so" if double else "liborc_f32.so")
        self.ft = C.c_double if double else C.c_float
        self.np_ft = np.float64 if double else np.float32
        L = self.lib
        assert L.orc_float_size() == (8 if double else 4)
        L.orc_profile_sample.restype = C.c_void_p
        L.orc_profile_sample.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_double]
        L.orc_profile_build.restype = C.c_void_p
        L.orc_profile_build.argtypes = [C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_profile_import.restype = C.c_void_p
        L.orc_profile_import.argtypes = [C.c_int, C.c_double] + [C.c_void_p] * 8
        L.orc_profile_del.argtypes = [C.c_void_p]
        L.orc_profile_M.argtypes = [C.c_void_p]
        for f in ("emM", "emI", "emN", "trans", "entry"):
            fn = getattr(L, "orc_profile_" + f)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p]
        L.orc_profile_ndist.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_sample_inputs.argtypes = [C.c_uint, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_specials.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_void_p]
        vit = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
               C.c_void_p, C.c_int]
        L.orc_viterbi_null.argtypes = vit
        L.orc_viterbi_alt.argtypes = vit
        L.orc_scores_fast.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_decode.argtypes = [C.c_void_p, C.c_uint, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p]
        L.orc_state_name.argtypes = [C.c_uint, C.c_char_p]
        L.orc_product_row.restype = C.c_long
        L.orc_product_row.argtypes = [C.c_void_p, C.c_long, C.c_long, C.c_char_p, C.c_double, C.c_double,
                                      C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_long]
        L.orc_scan.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                               C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]

    # ---- profiles -------------------------------------------------------
    def sample(self, seed, M, entry_dist=ENTRY_OCCUPANCY, eps=0.01):
        return OProfile(self, self.lib.orc_profile_sample(seed, M, entry_dist, float(eps)))

    def build(self, M, entry_dist, eps, null_lp, match_lp, trans):
        null_lp = np.ascontiguousarray(null_lp, np.float64)
        match_lp = np.ascontiguousarray(match_lp, np.float64)
        trans = np.ascontiguousarray(trans, np.float64)
        assert null_lp.size == 20 and match_lp.size == 20 * M and trans.size == 7 * (M + 1)
        return OProfile(self, self.lib.orc_profile_build(M, entry_dist, float(eps), null_lp.ctypes.data,
                                                         match_lp.ctypes.data, trans.ctypes.data))

    def import_tables(self, M, eps, emM, emI, emN, trans, entry, null_nd=None, ins_nd=None, match_nd=None):
        f = self.np_ft
        arrs = [np.ascontiguousarray(a, f) for a in (emM, emI, emN, trans, entry)]
        nds = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (null_nd, ins_nd, match_nd)]
        assert arrs[0].size == M * NTAB and arrs[3].size == 7 * (M + 1) and arrs[4].size == M
        ptrs = [a.ctypes.data for a in arrs] + [None if a is None else a.ctypes.data for a in nds]
        return OProfile(self, self.lib.orc_profile_import(M, float(eps), *ptrs))

    def set_threads(self, n):
        return self.lib.orc_set_threads(int(n))

    def sample_inputs(self, seed, M):
        nl = np.empty(20); ma = np.empty((M, 20)); tr = np.empty((M + 1, 7))
        self.lib.orc_sample_inputs(seed, M, nl.ctypes.data, ma.ctypes.data, tr.ctypes.data)
        return nl, ma, tr

    def specials(self, L, multi_hits=True, hmmer3_compat=False):
        out = np.empty(13, self.np_ft)
        rc = self.lib.orc_specials(L, int(multi_hits), int(hmmer3_compat), out.ctypes.data)
        return rc, out

    def state_name(self, sid):
        b = C.create_string_buffer(8)
        self.lib.orc_state_name(sid, b)
        return b.value.decode()

    def scan(self, profiles, seqs, multi_hits=True, hmmer3_compat=False, thr=10.0, flavour=1, want_paths=True):
        """thread_run restatement over all (seq, profile) pairs; returns dict of arrays in (seq, profile) order."""
        nprof, nseq = len(profiles), len(seqs)
        parr = (C.c_void_p * nprof)(*[p.h for p in profiles])
        bs = [s.encode() if isinstance(s, str) else s for s in seqs]
        sarr = (C.c_char_p * nseq)(*bs)
        lens = np.array([len(b) for b in bs], np.int32)
        n = nseq * nprof
        null = np.empty(n, self.np_ft); alt = np.empty(n, self.np_ft); hit = np.zeros(n, np.uint8)
        cap = int(sum(int(l) + 8 for l in lens)) * nprof if want_paths else 1
        off = np.zeros(n + 1, np.int32); st = np.zeros(cap, np.uint16); ln = np.zeros(cap, np.uint8)
        rc = self.lib.orc_scan(nprof, parr, nseq, sarr, lens.ctypes.data, int(multi_hits), int(hmmer3_compat),
                               float(thr), flavour, int(want_paths), null.ctypes.data, alt.ctypes.data,
                               hit.ctypes.data, off.ctypes.data if want_paths else None, st.ctypes.data,
                               ln.ctypes.data, cap)
        return dict(rc=rc, null=null.reshape(nseq, nprof), alt=alt.reshape(nseq, nprof),
                    hit=hit.reshape(nseq, nprof), path_off=off, step_state=st, step_len=ln)


class OProfile:
    def __init__(self, orc, h):
        assert h
        self.o, self.h = orc, h
        self.M = orc.lib.orc_profile_M(h)

    def __del__(self):
        try:
            self.o.lib.orc_profile_del(self.h)
        except Exception:
            pass

    def _arr(self, name, n):
        ptr = getattr(self.o.lib, "orc_profile_" + name)(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(self.o.ft)), shape=(n,)).copy()

    emM = property(lambda s: s._arr("emM", s.M * NTAB).reshape(s.M, NTAB))
    emI = property(lambda s: s._arr("emI", NTAB))
    emN = property(lambda s: s._arr("emN", NTAB))
    trans = property(lambda s: s._arr("trans", 7 * (s.M + 1)).reshape(s.M + 1, 7))
    entry = property(lambda s: s._arr("entry", s.M))

    def ndist(self, which):
        out = np.empty(129)
        self.o.lib.orc_profile_ndist(self.h, which, out.ctypes.data)
        return out

    def _vit(self, fn, seq, multi_hits, hmmer3_compat, want_path=True):
        b = seq.encode() if isinstance(seq, str) else seq
        L = len(b)
        ll = self.o.ft()
        cap = L + 8
        st = np.zeros(cap, np.uint16); ln = np.zeros(cap, np.uint8); ns = C.c_int(0)
        rc = fn(self.h, b, L, int(multi_hits), int(hmmer3_compat), C.byref(ll),
                st.ctypes.data if want_path else None, ln.ctypes.data if want_path else None,
                C.byref(ns) if want_path else None, cap)
        return rc, ll.value, list(zip(st[:ns.value].tolist(), ln[:ns.value].tolist()))

    def viterbi_null(self, seq, multi_hits=True, hmmer3_compat=False):
        return self._vit(self.o.lib.orc_viterbi_null, seq, multi_hits, hmmer3_compat)

    def viterbi_alt(self, seq, multi_hits=True, hmmer3_compat=False):
        return self._vit(self.o.lib.orc_viterbi_alt, seq, multi_hits, hmmer3_compat)

    def scores_fast(self, seq, multi_hits=True, hmmer3_compat=False):
        b = seq.encode() if isinstance(seq, str) else seq
        n, a = self.o.ft(), self.o.ft()
        rc = self.o.lib.orc_scores_fast(self.h, b, len(b), int(multi_hits), int(hmmer3_compat), C.byref(n), C.byref(a))
        return rc, n.value, a.value

    def decode(self, state_id, frag):
        b = frag.encode() if isinstance(frag, str) else frag
        cod = C.create_string_buffer(4); am = C.create_string_buffer(2)
        rc = self.o.lib.orc_decode(self.h, state_id, b, len(b), cod, am)
        return rc, cod.raw[:3].decode(), am.raw[:1].decode()

    def product_row(self, scan_id, seq_id, accession, alt, null, seq, path):
        b = seq.encode() if isinstance(seq, str) else seq
        st = np.array([p[0] for p in path], np.uint16); ln = np.array([p[1] for p in path], np.uint8)
        cap = 64 * (len(path) + 8) + 256
        buf = C.create_string_buffer(cap)
        n = self.o.lib.orc_product_row(self.h, scan_id, seq_id, accession.encode(), float(alt), float(null), b,
                                       st.ctypes.data, ln.ctypes.data, len(path), buf, cap)
        assert n >= 0
        return buf.raw[:n].decode()
