"""deciphon-old_b200 -- ctypes view of libdcpgpu.so (the C ABI in include/dcpgpu.h).

The product is the C/CUDA library; this module only loads it and gives tests and bench.py
thin Python handles with the reference's names (protein_profile_sample, protein_profile_setup,
protein_profile_decode, thread_run's scan loop, prod_fwrite).  There is no Python or CPU
implementation of the scan here: without the library, or without a CUDA device, everything
below raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DCPGPU_LIB") or os.path.join(_HERE, "libdcpgpu.so")  # override: kernel-variant experiments

RC_OK, RC_END, RC_EFAIL, RC_EINVAL, RC_EIO, RC_ENOMEM, RC_EPARSE, RC_EAPI, RC_EHTTP = range(9)
ENTRY_DIST_NULL, ENTRY_DIST_UNIFORM, ENTRY_DIST_OCCUPANCY = range(3)
FRAME_TABLE_SIZE = 1364
PROTEIN_R_STATE, PROTEIN_S_STATE, PROTEIN_N_STATE, PROTEIN_B_STATE = 0xC000, 0xC001, 0xC002, 0xC003
PROTEIN_E_STATE, PROTEIN_J_STATE, PROTEIN_C_STATE, PROTEIN_T_STATE = 0xC004, 0xC005, 0xC006, 0xC007

# every symbol include/dcpgpu.h declares
EXPORTS = [
    "protein_model_new", "protein_model_setup", "protein_model_add_node", "protein_model_add_trans",
    "protein_model_del", "protein_profile_new", "protein_profile_absorb", "protein_profile_sample", "protein_profile_build",
    "protein_profile_setup", "protein_profile_decode", "protein_profile_del", "protein_profile_core_size",
    "protein_profile_accession", "protein_profile_match_emission", "protein_profile_insert_emission",
    "protein_profile_null_emission", "protein_profile_trans", "protein_profile_entry",
    "protein_profile_nuclt_dist", "protein_state_name", "protein_state_is_mute", "xmath_lrt_f32",
    "protein_h3reader_new", "protein_h3reader_next", "protein_h3reader_model", "protein_h3reader_accession",
    "protein_h3reader_name", "protein_h3reader_del", "protein_db_writer_open", "protein_db_writer_pack_profile",
    "protein_db_writer_close", "protein_db_reader_open", "protein_db_reader_nprofiles", "protein_db_reader_cfg",
    "protein_db_reader_profile_size", "protein_db_reader_next", "protein_db_reader_close", "dcpgpu_press_hmm", "dcpgpu_db_accession", "dcpgpu_db_core_size",
    "dcpgpu_db_new", "dcpgpu_db_add", "dcpgpu_db_commit", "dcpgpu_db_nprofiles", "dcpgpu_db_device_bytes",
    "dcpgpu_db_del", "dcpgpu_seqs_new", "dcpgpu_seqs_del", "dcpgpu_scan_resident", "dcpgpu_scan",
    "dcpgpu_result_nseqs", "dcpgpu_result_nprofiles", "dcpgpu_result_null_loglik", "dcpgpu_result_alt_loglik",
    "dcpgpu_result_hit", "dcpgpu_result_nhits", "dcpgpu_result_hit_at", "dcpgpu_result_hits", "dcpgpu_result_steps",
    "dcpgpu_result_timing",
    "dcpgpu_result_del", "dcpgpu_shard_profiles", "dcpgpu_kernel_shape", "dcpgpu_kernel_padded_width", "dcpgpu_profile_cost", "dcpgpu_shard_sequences",
    "dcpgpu_mdb_new", "dcpgpu_mdb_add", "dcpgpu_mdb_commit", "dcpgpu_mdb_view", "dcpgpu_mdb_ndevices",
    "dcpgpu_mdb_nprofiles", "dcpgpu_mdb_axis", "dcpgpu_mdb_imbalance", "dcpgpu_mdb_device_of", "dcpgpu_mdb_device_bytes",
    "dcpgpu_mdb_scan", "dcpgpu_mdb_del", "dcpgpu_result_nparts", "dcpgpu_result_part_timing", "dcpgpu_prod_fwrite_header", "dcpgpu_prod_fwrite",
    "dcpgpu_prod_row", "dcpgpu_microbench_alu", "dcpgpu_last_error",
]


class DcpError(RuntimeError):
    def __init__(self, rc, msg):
        super().__init__("rc=%d: %s" % (rc, msg))
        self.rc = rc


class _Cfg(C.Structure):
    _fields_ = [("entry_dist", C.c_int), ("epsilon", C.c_float)]


class _Trans(C.Structure):
    _fields_ = [("data", C.c_float * 7)]


PROGRESS_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64)


class _Params(C.Structure):
    _fields_ = [("multi_hits", C.c_bool), ("hmmer3_compat", C.c_bool), ("lrt_threshold", C.c_double),
                ("want_paths", C.c_bool), ("progress", PROGRESS_FN), ("user", C.c_void_p)]


class _Step(C.Structure):
    _fields_ = [("state_id", C.c_uint16), ("seqlen", C.c_uint8)]


class Timing(C.Structure):
    _fields_ = [("prep_ms", C.c_float), ("score_ms", C.c_float), ("trace_ms", C.c_float), ("total_ms", C.c_float),
                ("launches", C.c_uint64), ("alt_cells", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64)]


_lib = None


def lib():
    """Load libdcpgpu.so (built in-tree by __graft_entry__.build()).  Fails loudly if missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: run __graft_entry__.build() (make -C deciphon-old_b200/csrc)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u, i = C.c_void_p, C.c_uint, C.c_int
    L.protein_model_new.restype = vp
    L.protein_model_new.argtypes = [_Cfg, vp]
    L.protein_model_setup.argtypes = [vp, u]
    L.protein_model_add_node.argtypes = [vp, vp, C.c_char]
    L.protein_model_add_trans.argtypes = [vp, _Trans]
    L.protein_model_del.argtypes = [vp]
    L.protein_profile_new.restype = vp
    L.protein_profile_new.argtypes = [C.c_char_p, _Cfg]
    L.protein_profile_absorb.argtypes = [vp, vp]
    L.protein_profile_sample.argtypes = [vp, u, u]
    L.protein_profile_build.argtypes = [vp, u, vp, vp, vp, C.c_char_p]
    L.protein_profile_setup.argtypes = [vp, u, C.c_bool, C.c_bool, vp]
    L.protein_profile_decode.argtypes = [vp, C.c_char_p, u, u, C.c_char_p, C.c_char_p]
    L.protein_profile_del.argtypes = [vp]
    L.protein_profile_core_size.argtypes = [vp]
    L.protein_profile_accession.restype = C.c_char_p
    L.protein_profile_accession.argtypes = [vp]
    for f in ("match_emission", "insert_emission", "null_emission", "trans", "entry"):
        fn = getattr(L, "protein_profile_" + f)
        fn.restype = C.POINTER(C.c_float)
        fn.argtypes = [vp]
    L.protein_profile_nuclt_dist.argtypes = [vp, i, vp]
    L.protein_state_name.argtypes = [u, C.c_char_p]
    L.protein_state_is_mute.restype = C.c_bool
    L.protein_state_is_mute.argtypes = [u]
    L.xmath_lrt_f32.restype = C.c_float
    L.xmath_lrt_f32.argtypes = [C.c_float, C.c_float]
    L.dcpgpu_db_new.argtypes = [C.POINTER(vp), i]
    L.dcpgpu_db_add.argtypes = [vp, vp]
    L.dcpgpu_db_commit.argtypes = [vp]
    L.dcpgpu_db_nprofiles.argtypes = [vp]
    L.dcpgpu_db_device_bytes.restype = C.c_uint64
    L.dcpgpu_db_device_bytes.argtypes = [vp]
    L.dcpgpu_db_del.argtypes = [vp]
    L.dcpgpu_seqs_new.argtypes = [C.POINTER(vp), vp, u, vp, vp]
    L.dcpgpu_seqs_del.argtypes = [vp]
    L.dcpgpu_scan_resident.argtypes = [vp, vp, C.POINTER(_Params), C.POINTER(vp)]
    L.dcpgpu_scan.argtypes = [vp, u, vp, vp, C.POINTER(_Params), C.POINTER(vp)]
    L.dcpgpu_result_nseqs.argtypes = [vp]
    L.dcpgpu_result_nprofiles.argtypes = [vp]
    for f in ("null_loglik", "alt_loglik"):
        fn = getattr(L, "dcpgpu_result_" + f)
        fn.restype = C.POINTER(C.c_float)
        fn.argtypes = [vp]
    L.dcpgpu_result_hit.restype = C.POINTER(C.c_uint8)
    L.dcpgpu_result_hit.argtypes = [vp]
    L.dcpgpu_result_nhits.restype = C.c_uint64
    L.dcpgpu_result_nhits.argtypes = [vp]
    L.dcpgpu_result_hit_at.argtypes = [vp, C.c_uint64, C.POINTER(u), C.POINTER(u), C.POINTER(C.POINTER(_Step)),
                                       C.POINTER(u)]
    L.dcpgpu_result_hits.argtypes = [vp, vp, vp, vp, vp, vp]
    L.dcpgpu_result_steps.restype = C.c_uint64
    L.dcpgpu_result_steps.argtypes = [vp, C.POINTER(C.POINTER(_Step))]
    L.dcpgpu_result_timing.argtypes = [vp, C.POINTER(Timing)]
    L.dcpgpu_result_del.argtypes = [vp]
    L.dcpgpu_shard_profiles.argtypes = [u, vp, u, vp]
    L.dcpgpu_kernel_shape.argtypes = [u, vp, vp, vp]
    L.dcpgpu_kernel_shape.restype = C.c_int
    L.dcpgpu_kernel_padded_width.argtypes = [u]
    L.dcpgpu_kernel_padded_width.restype = u
    L.dcpgpu_profile_cost.argtypes = [u]
    L.dcpgpu_profile_cost.restype = C.c_double
    L.dcpgpu_shard_sequences.argtypes = [u, vp, u, vp]
    L.dcpgpu_mdb_new.argtypes = [C.POINTER(vp), u, vp]
    L.dcpgpu_mdb_add.argtypes = [vp, vp]
    L.dcpgpu_mdb_commit.argtypes = [vp, i]
    L.dcpgpu_mdb_view.restype = vp
    L.dcpgpu_mdb_view.argtypes = [vp]
    L.dcpgpu_mdb_ndevices.argtypes = [vp]
    L.dcpgpu_mdb_nprofiles.argtypes = [vp]
    L.dcpgpu_mdb_axis.argtypes = [vp]
    L.dcpgpu_mdb_imbalance.restype = C.c_double
    L.dcpgpu_mdb_imbalance.argtypes = [vp]
    L.dcpgpu_mdb_device_of.argtypes = [vp, u]
    L.dcpgpu_mdb_device_of.restype = i
    L.dcpgpu_mdb_device_bytes.restype = C.c_uint64
    L.dcpgpu_mdb_device_bytes.argtypes = [vp, u]
    L.dcpgpu_mdb_scan.argtypes = [vp, u, vp, vp, C.POINTER(_Params), C.POINTER(vp)]
    L.dcpgpu_mdb_del.argtypes = [vp]
    L.dcpgpu_result_nparts.argtypes = [vp]
    L.dcpgpu_result_part_timing.argtypes = [vp, u, C.POINTER(i), C.POINTER(Timing)]
    L.dcpgpu_prod_row.restype = C.c_long
    L.dcpgpu_prod_row.argtypes = [vp, vp, C.c_uint64, C.c_int64, C.c_int64, C.c_char_p, C.c_char_p, C.c_long]
    L.dcpgpu_microbench_alu.argtypes = [i, vp]
    L.protein_h3reader_new.restype = vp
    L.protein_h3reader_new.argtypes = [_Cfg, vp]
    L.protein_h3reader_next.argtypes = [vp]
    L.protein_h3reader_model.restype = vp
    L.protein_h3reader_model.argtypes = [vp]
    L.protein_h3reader_accession.restype = C.c_char_p
    L.protein_h3reader_accession.argtypes = [vp]
    L.protein_h3reader_name.restype = C.c_char_p
    L.protein_h3reader_name.argtypes = [vp]
    L.protein_h3reader_del.argtypes = [vp]
    L.dcpgpu_press_hmm.argtypes = [vp, vp, _Cfg, C.POINTER(u)]
    L.protein_db_writer_open.restype = vp
    L.protein_db_writer_open.argtypes = [vp, _Cfg]
    L.protein_db_writer_pack_profile.argtypes = [vp, vp]
    L.protein_db_writer_close.argtypes = [vp, C.c_bool]
    L.protein_db_reader_open.argtypes = [C.POINTER(vp), vp]
    L.protein_db_reader_nprofiles.argtypes = [vp]
    L.protein_db_reader_cfg.restype = _Cfg
    L.protein_db_reader_cfg.argtypes = [vp]
    L.protein_db_reader_profile_size.restype = C.c_uint32
    L.protein_db_reader_profile_size.argtypes = [vp, u]
    L.protein_db_reader_next.argtypes = [vp, C.POINTER(vp)]
    L.protein_db_reader_close.argtypes = [vp]
    L.dcpgpu_db_accession.restype = C.c_char_p
    L.dcpgpu_db_accession.argtypes = [vp, u]
    L.dcpgpu_db_core_size.argtypes = [vp, u]
    L.dcpgpu_last_error.restype = C.c_char_p
    _lib = L
    return L


def _check(rc):
    if rc != RC_OK:
        raise DcpError(rc, lib().dcpgpu_last_error().decode())


def protein_cfg(entry_dist=ENTRY_DIST_OCCUPANCY, epsilon=0.01):
    return _Cfg(entry_dist, float(epsilon))


def protein_state_name(state_id):
    buf = C.create_string_buffer(8)
    lib().protein_state_name(state_id, buf)
    return buf.value.decode()


def protein_state_is_mute(state_id):
    return bool(lib().protein_state_is_mute(state_id))


def xmath_lrt(null, alt):
    return lib().xmath_lrt_f32(null, alt)


def microbench_alu(device=0):
    """Measured FP32 issue peaks: dict of 1e9 lane-instructions/s for FADD, FMNMX3 and the DP-cell mix."""
    out = np.zeros(6)
    _check(lib().dcpgpu_microbench_alu(device, out.ctypes.data))
    return {"fadd_ginst": out[0], "fmnmx3_ginst": out[1], "mix_ginst": out[2], "sms": int(out[3]),
            "mix_clock_mhz": out[4], "fadd_clock_mhz": out[5]}


def kernel_shape(core_size):
    """(warps per pair, nodes per lane, blocks per pair) the engine uses for a profile of `core_size` nodes."""
    w, q, b = C.c_uint(), C.c_uint(), C.c_uint()
    _check(lib().dcpgpu_kernel_shape(core_size, C.byref(w), C.byref(q), C.byref(b)))
    return w.value, q.value, b.value


def kernel_padded_width(core_size):
    """Nodes the kernels compute for a profile of `core_size` nodes (its kernel class's capacity)."""
    return int(lib().dcpgpu_kernel_padded_width(int(core_size)))


def shard_profiles(core_sizes, nshards):
    """Device index of every profile: longest-processing-time partition by modelled cost."""
    cs = np.ascontiguousarray(core_sizes, np.uint32)
    out = np.zeros(len(cs), np.uint32)
    _check(lib().dcpgpu_shard_profiles(len(cs), cs.ctypes.data, nshards, out.ctypes.data))
    return out


def profile_cost(core_size):
    """Modelled score-pass time of one sequence row against a profile of `core_size` nodes (ns on one B200)."""
    return float(lib().dcpgpu_profile_cost(int(core_size)))


def shard_sequences(lens, nshards):
    """nshards + 1 bounds of contiguous sequence ranges with about equal nucleotide totals."""
    ln = np.ascontiguousarray(lens, np.uint32)
    out = np.zeros(nshards + 1, np.uint32)
    _check(lib().dcpgpu_shard_sequences(len(ln), ln.ctypes.data, nshards, out.ctypes.data))
    return out


class ProteinProfile:
    """struct protein_profile (include/deciphon/model/protein_profile.h:12-43)."""

    def __init__(self, accession="accession", cfg=None):
        cfg = cfg or protein_cfg()
        self.h = lib().protein_profile_new(accession.encode(), cfg)
        if not self.h:
            raise DcpError(RC_EINVAL, lib().dcpgpu_last_error().decode())

    def __del__(self):
        try:
            if self.h:
                lib().protein_profile_del(self.h)
        except Exception:
            pass

    @classmethod
    def sample(cls, seed, core_size, cfg=None, accession="accession"):
        p = cls(accession, cfg)
        _check(lib().protein_profile_sample(p.h, seed, core_size))
        return p

    @classmethod
    def from_model(cls, null_lprobs, match_lprobs, trans, cfg=None, accession="accession", consensus=None):
        """protein_model_init/setup/add_node/add_trans + protein_profile_absorb, as protein_h3reader_next does."""
        L = lib()
        cfg = cfg or protein_cfg()
        nl = np.ascontiguousarray(null_lprobs, np.float32)
        ml = np.ascontiguousarray(match_lprobs, np.float32)
        tr = np.ascontiguousarray(trans, np.float32)
        M = ml.shape[0]
        assert nl.shape == (20,) and ml.shape == (M, 20) and tr.shape == (M + 1, 7)
        m = L.protein_model_new(cfg, nl.ctypes.data)
        if not m:
            raise DcpError(RC_EINVAL, L.dcpgpu_last_error().decode())
        try:
            _check(L.protein_model_setup(m, M))
            _check(L.protein_model_add_trans(m, _Trans((C.c_float * 7)(*tr[0]))))
            for k in range(M):
                c = (consensus[k] if consensus else "-").encode()
                _check(L.protein_model_add_node(m, ml[k].ctypes.data, c))
                _check(L.protein_model_add_trans(m, _Trans((C.c_float * 7)(*tr[k + 1]))))
            p = cls(accession, cfg)
            _check(L.protein_profile_absorb(p.h, m))
        finally:
            L.protein_model_del(m)
        return p

    @classmethod
    def build(cls, null_lprobs, match_lprobs, trans, cfg=None, accession="accession", consensus=None):
        """Same as from_model in one C call (releases the GIL; used to build large synthetic databases)."""
        nl = np.ascontiguousarray(null_lprobs, np.float32)
        ml = np.ascontiguousarray(match_lprobs, np.float32)
        tr = np.ascontiguousarray(trans, np.float32)
        M = ml.shape[0]
        assert nl.shape == (20,) and ml.shape == (M, 20) and tr.shape == (M + 1, 7)
        p = cls(accession, cfg)
        _check(lib().protein_profile_build(p.h, M, nl.ctypes.data, ml.ctypes.data, tr.ctypes.data,
                                           consensus.encode() if consensus else None))
        return p

    core_size = property(lambda s: lib().protein_profile_core_size(s.h))
    accession = property(lambda s: lib().protein_profile_accession(s.h).decode())

    def _arr(self, name, n):
        ptr = getattr(lib(), "protein_profile_" + name)(self.h)
        return np.ctypeslib.as_array(ptr, shape=(n,)).copy()

    match_emission = property(lambda s: s._arr("match_emission", s.core_size * FRAME_TABLE_SIZE)
                              .reshape(s.core_size, FRAME_TABLE_SIZE))
    insert_emission = property(lambda s: s._arr("insert_emission", FRAME_TABLE_SIZE))
    null_emission = property(lambda s: s._arr("null_emission", FRAME_TABLE_SIZE))
    trans = property(lambda s: s._arr("trans", 7 * (s.core_size + 1)).reshape(s.core_size + 1, 7))
    entry = property(lambda s: s._arr("entry", s.core_size))

    def nuclt_dist(self, which):
        out = np.empty(129)
        _check(lib().protein_profile_nuclt_dist(self.h, which, out.ctypes.data))
        return out

    def setup(self, seq_size, multi_hits=True, hmmer3_compat=False):
        """protein_profile_setup; returns (rc, 13 special transition scores)."""
        out = np.zeros(13, np.float32)
        rc = lib().protein_profile_setup(self.h, seq_size, multi_hits, hmmer3_compat, out.ctypes.data)
        return rc, out

    def decode(self, frag, state_id):
        b = frag.encode() if isinstance(frag, str) else frag
        cod = C.create_string_buffer(4)
        am = C.create_string_buffer(2)
        rc = lib().protein_profile_decode(self.h, b, len(b), state_id, cod, am)
        return rc, cod.raw[:3].decode(errors="replace"), am.raw[:1].decode(errors="replace")


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


def read_hmm(path, cfg=None):
    """protein_h3reader_next + protein_profile_absorb over a HMMER3 ASCII file -> [ProteinProfile]."""
    L = lib()
    cfg = cfg or protein_cfg()
    fp = _libc.fopen(os.fsencode(path), b"r")
    if not fp:
        raise DcpError(RC_EIO, "cannot open %s" % path)
    rd = L.protein_h3reader_new(cfg, fp)
    out = []
    try:
        while True:
            rc = L.protein_h3reader_next(rd)
            if rc == RC_END:
                break
            _check(rc)
            p = ProteinProfile(L.protein_h3reader_accession(rd).decode(), cfg)
            _check(L.protein_profile_absorb(p.h, L.protein_h3reader_model(rd)))
            out.append(p)
    finally:
        L.protein_h3reader_del(rd)
        _libc.fclose(fp)
    return out


def write_dcp(path, profiles, cfg=None):
    """protein_db_writer_open / _pack_profile / db_writer_close."""
    L = lib()
    cfg = cfg or protein_cfg()
    fp = _libc.fopen(os.fsencode(path), b"wb")
    if not fp:
        raise DcpError(RC_EIO, "cannot open %s" % path)
    w = L.protein_db_writer_open(fp, cfg)
    ok = False
    try:
        for p in profiles:
            _check(L.protein_db_writer_pack_profile(w, p.h))
        ok = True
    finally:
        rc = L.protein_db_writer_close(w, ok)
        _libc.fclose(fp)
    _check(rc)


def read_dcp(path):
    """protein_db_reader_open + profile_reader_next -> (cfg, [ProteinProfile])."""
    L = lib()
    fp = _libc.fopen(os.fsencode(path), b"rb")
    if not fp:
        raise DcpError(RC_EIO, "cannot open %s" % path)
    rd = C.c_void_p()
    try:
        _check(L.protein_db_reader_open(C.byref(rd), fp))
        cfg = L.protein_db_reader_cfg(rd)
        out = []
        while True:
            h = C.c_void_p()
            rc = L.protein_db_reader_next(rd, C.byref(h))
            if rc == RC_END:
                break
            _check(rc)
            p = ProteinProfile.__new__(ProteinProfile)
            p.h = h
            out.append(p)
        assert len(out) == L.protein_db_reader_nprofiles(rd)
    finally:
        if rd:
            L.protein_db_reader_close(rd)
        _libc.fclose(fp)
    return cfg, out


def _seq_arrays(seqs):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    arr = (C.c_char_p * len(bs))(*bs)
    lens = np.array([len(b) for b in bs], np.uint32)
    return bs, arr, lens


class Seqs:
    def __init__(self, db, seqs):
        self.db = db
        self.bytes, arr, lens = _seq_arrays(seqs)
        self.h = C.c_void_p()
        _check(lib().dcpgpu_seqs_new(C.byref(self.h), db.h, len(self.bytes), arr, lens.ctypes.data))

    def __del__(self):
        try:
            if self.h:
                lib().dcpgpu_seqs_del(self.h)
        except Exception:
            pass


class Result:
    def __init__(self, db, h, seqs_bytes):
        self.db, self.h, self.seqs = db, h, seqs_bytes
        self.nseqs = lib().dcpgpu_result_nseqs(h)
        self.nprofiles = lib().dcpgpu_result_nprofiles(h)

    def __del__(self):
        try:
            if self.h:
                lib().dcpgpu_result_del(self.h)
        except Exception:
            pass

    def _mat(self, fn, dt):
        ptr = fn(self.h)
        if not ptr:
            raise DcpError(RC_EFAIL, lib().dcpgpu_last_error().decode())
        return np.ctypeslib.as_array(ptr, shape=(self.nseqs * self.nprofiles,)).astype(dt).reshape(
            self.nseqs, self.nprofiles)

    null_loglik = property(lambda s: s._mat(lib().dcpgpu_result_null_loglik, np.float32))
    alt_loglik = property(lambda s: s._mat(lib().dcpgpu_result_alt_loglik, np.float32))
    hit = property(lambda s: s._mat(lib().dcpgpu_result_hit, np.uint8))
    nhits = property(lambda s: int(lib().dcpgpu_result_nhits(s.h)))

    def hit_at(self, i):
        """-> (seq_idx, prof_idx, [(state_id, seqlen), ...])"""
        si, pi, n = C.c_uint(), C.c_uint(), C.c_uint()
        steps = C.POINTER(_Step)()
        _check(lib().dcpgpu_result_hit_at(self.h, i, C.byref(si), C.byref(pi), C.byref(steps), C.byref(n)))
        path = [(steps[k].state_id, steps[k].seqlen) for k in range(n.value)] if steps else []
        return si.value, pi.value, path

    def hits(self):
        """The hit list as arrays: (seq u32, prof u32, alt f32, null f32, nsteps u32), in (sequence, profile) order."""
        n = self.nhits
        seq, prof, ns = np.zeros(n, np.uint32), np.zeros(n, np.uint32), np.zeros(n, np.uint32)
        alt, null = np.zeros(n, np.float32), np.zeros(n, np.float32)
        _check(lib().dcpgpu_result_hits(self.h, seq.ctypes.data, prof.ctypes.data, alt.ctypes.data, null.ctypes.data,
                                        ns.ctypes.data))
        return seq, prof, alt, null, ns

    def steps(self):
        """All paths back to back as an (nsteps_total, 2) uint16 array of (state_id, seqlen)."""
        ptr = C.POINTER(_Step)()
        n = int(lib().dcpgpu_result_steps(self.h, C.byref(ptr)))
        if n == 0:
            return np.zeros((0, 2), np.uint16)
        raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint16)), shape=(n, 2)).copy()
        raw[:, 1] &= 0xFF  # seqlen is one byte; the other is struct padding
        return raw

    def hit_scores(self, i):
        s = self.hits()
        return float(s[2][i]), float(s[3][i])

    def product_row(self, i, scan_id=0, seq_id=None):
        si, _, path = self.hit_at(i)
        cap = 64 * (len(path) + 8) + 512
        buf = C.create_string_buffer(cap)
        n = lib().dcpgpu_prod_row(self.h, self.db.h, i, scan_id, si if seq_id is None else seq_id, self.seqs[si], buf,
                                  cap)
        if n < 0:
            raise DcpError(RC_EIO, lib().dcpgpu_last_error().decode())
        return buf.raw[:n].decode()

    @property
    def timing(self):
        t = Timing()
        lib().dcpgpu_result_timing(self.h, C.byref(t))
        return t

    @property
    def part_timings(self):
        """[(device, Timing)] of the launch sets a merged (tiled or multi-device) result was built from."""
        out = []
        for k in range(lib().dcpgpu_result_nparts(self.h)):
            t, dev = Timing(), C.c_int()
            _check(lib().dcpgpu_result_part_timing(self.h, k, C.byref(dev), C.byref(t)))
            out.append((dev.value, t))
        return out


class Db:
    """Profiles resident in HBM on one device (replaces profile_reader + per-pair unpack)."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().dcpgpu_db_new(C.byref(self.h), device))
        self.profiles = []

    def __del__(self):
        try:
            if self.h:
                lib().dcpgpu_db_del(self.h)
        except Exception:
            pass

    def add(self, prof):
        _check(lib().dcpgpu_db_add(self.h, prof.h))
        self.profiles.append(prof)

    def press(self, hmm_path, cfg=None):
        """hmm_press: add every profile of a HMMER3 file (call commit afterwards)."""
        fp = _libc.fopen(os.fsencode(hmm_path), b"r")
        if not fp:
            raise DcpError(RC_EIO, "cannot open %s" % hmm_path)
        n = C.c_uint(0)
        try:
            _check(lib().dcpgpu_press_hmm(self.h, fp, cfg or protein_cfg(), C.byref(n)))
        finally:
            _libc.fclose(fp)
        return n.value

    def accession(self, i):
        return lib().dcpgpu_db_accession(self.h, i).decode()

    def commit(self):
        _check(lib().dcpgpu_db_commit(self.h))

    nprofiles = property(lambda s: lib().dcpgpu_db_nprofiles(s.h))
    device_bytes = property(lambda s: int(lib().dcpgpu_db_device_bytes(s.h)))

    @staticmethod
    def _params(multi_hits, hmmer3_compat, lrt_threshold, want_paths, progress=None):
        cb = PROGRESS_FN(lambda user, pairs: progress(int(pairs))) if progress else PROGRESS_FN()
        return _Params(multi_hits, hmmer3_compat, float(lrt_threshold), want_paths, cb, None)

    def stage(self, seqs):
        return Seqs(self, seqs)

    def scan_resident(self, staged, multi_hits=True, hmmer3_compat=False, lrt_threshold=10.0, want_paths=True):
        prm = self._params(multi_hits, hmmer3_compat, lrt_threshold, want_paths)
        out = C.c_void_p()
        _check(lib().dcpgpu_scan_resident(self.h, staged.h, C.byref(prm), C.byref(out)))
        return Result(self, out, staged.bytes)

    def scan(self, seqs, multi_hits=True, hmmer3_compat=False, lrt_threshold=10.0, want_paths=True, progress=None):
        """thread_run's loop for every (sequence, profile) pair, from host buffers."""
        bs, arr, lens = _seq_arrays(seqs)
        prm = self._params(multi_hits, hmmer3_compat, lrt_threshold, want_paths, progress)
        out = C.c_void_p()
        _check(lib().dcpgpu_scan(self.h, len(bs), arr, lens.ctypes.data, C.byref(prm), C.byref(out)))
        return Result(self, out, bs)


AXIS_AUTO, AXIS_PROFILES, AXIS_SEQUENCES = range(3)


class _View:
    """dcpgpu_mdb_view: the global profile list (what product rows are written against)."""

    def __init__(self, h):
        self.h = h


class Mdb:
    """A database over several GPUs of one box (replaces scan_run's omp partitions, scan.c:239-250)."""

    def __init__(self, devices):
        dv = np.ascontiguousarray(devices, np.int32)
        self.h = C.c_void_p()
        _check(lib().dcpgpu_mdb_new(C.byref(self.h), len(dv), dv.ctypes.data))
        self.view = _View(lib().dcpgpu_mdb_view(self.h))

    def __del__(self):
        try:
            if self.h:
                lib().dcpgpu_mdb_del(self.h)
        except Exception:
            pass

    def add(self, prof):
        _check(lib().dcpgpu_mdb_add(self.h, prof.h))

    def commit(self, axis=AXIS_AUTO):
        _check(lib().dcpgpu_mdb_commit(self.h, axis))

    ndevices = property(lambda s: lib().dcpgpu_mdb_ndevices(s.h))
    nprofiles = property(lambda s: lib().dcpgpu_mdb_nprofiles(s.h))
    axis = property(lambda s: lib().dcpgpu_mdb_axis(s.h))
    imbalance = property(lambda s: lib().dcpgpu_mdb_imbalance(s.h))

    def device_of(self, profile):
        return lib().dcpgpu_mdb_device_of(self.h, profile)

    def scan(self, seqs, multi_hits=True, hmmer3_compat=False, lrt_threshold=10.0, want_paths=True, progress=None):
        bs, arr, lens = _seq_arrays(seqs)
        prm = Db._params(multi_hits, hmmer3_compat, lrt_threshold, want_paths, progress)
        out = C.c_void_p()
        _check(lib().dcpgpu_mdb_scan(self.h, len(bs), arr, lens.ctypes.data, C.byref(prm), C.byref(out)))
        r = Result(self.view, out, bs)
        r._keep = self  # the merged result refers to the mdb's shard maps
        return r
