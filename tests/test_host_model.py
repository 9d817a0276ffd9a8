"""Host model layer of libdcpgpu.so (C) against the oracle: tables, entry distribution, specials,
decode, state names, the C-ABI export list.  No GPU needed."""
import ctypes as C

import numpy as np
import pytest

import orc
from common import GOLD, SEQ32, oracle_twin, plan7_profile_inputs

EPS_01F = float(np.float32(0.1))


def ulp_diff(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.lib()
    import re, os
    hdr = open(os.path.join(os.path.dirname(pkg.__file__), "..", "include", "dcpgpu.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b((?:dcpgpu|protein|xmath)_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"protein_cfg"}
    assert declared == set(pkg.EXPORTS)
    for s in declared:
        assert hasattr(lib, s), s


@pytest.mark.parametrize("entry", [1, 2])
@pytest.mark.parametrize("seed,M,eps", [(1, 2, EPS_01F), (2, 5, 0.01), (3, 17, 0.01), (4, 64, EPS_01F)])
def test_sampled_profile_tables_match_oracle(pkg, o32, seed, M, eps, entry):
    p = pkg.ProteinProfile.sample(seed, M, pkg.protein_cfg(entry, eps))
    q = o32.sample(seed, M, entry, float(np.float32(eps)))
    assert p.core_size == M
    assert np.array_equal(p.trans, q.trans)
    # tables: product sums probabilities, oracle sums in the log domain; both round once to fp32
    for a, b in ((p.match_emission, q.emM), (p.insert_emission, q.emI), (p.null_emission, q.emN)):
        fin = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), fin)
        d = ulp_diff(a[fin], b[fin])
        assert d.max() <= 1 and (d > 0).mean() < 1e-3
    d = ulp_diff(p.entry, q.entry)
    assert d.max() <= 1
    for which in (-2, -1, 0, M - 1):
        assert np.allclose(p.nuclt_dist(which), q.ndist(which), rtol=1e-12, atol=1e-12)


def test_model_api_equals_sampler_and_oracle_build(pkg, o32):
    nl, ma, tr = o32.sample_inputs(9, 12)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    a = pkg.ProteinProfile.from_model(nl, ma, tr, cfg)
    b = pkg.ProteinProfile.sample(9, 12, cfg)
    assert np.array_equal(a.match_emission, b.match_emission) and np.array_equal(a.entry, b.entry)
    q = o32.build(12, 2, float(np.float32(0.01)), nl, ma, tr)
    assert ulp_diff(a.match_emission, q.emM).max() <= 1


def test_model_error_paths(pkg):
    L = pkg.lib()
    cfg = pkg.protein_cfg()
    nl = np.zeros(20, np.float32)
    m = L.protein_model_new(cfg, nl.ctypes.data)
    assert L.protein_model_add_node(m, nl.ctypes.data, b"-") == pkg.RC_EFAIL  # setup not called
    assert L.protein_model_setup(m, 0) == pkg.RC_EINVAL
    assert L.protein_model_setup(m, 4097) == pkg.RC_EINVAL
    assert L.protein_model_setup(m, 1) == pkg.RC_OK
    assert L.protein_model_add_node(m, nl.ctypes.data, b"-") == pkg.RC_OK
    assert L.protein_model_add_node(m, nl.ctypes.data, b"-") == pkg.RC_EFAIL  # reached limit of nodes
    L.protein_model_del(m)


def test_profile_setup_specials(pkg, o32):
    p = pkg.ProteinProfile.sample(1, 2)
    rc, _ = p.setup(0)
    assert rc == pkg.RC_EINVAL  # test/protein_profile.c:31
    for L in (1, 2, 32, 150, 1000, 1053, 10000, 100003):
        for mh in (True, False):
            for h3 in (True, False):
                rc, x = p.setup(L, mh, h3)
                rc2, y = o32.specials(L, mh, h3)
                assert rc == 0 and rc2 == 0
                assert np.array_equal(x, y, equal_nan=True), (L, mh, h3, x, y)


def test_decode_and_names(pkg, o32):
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_UNIFORM, EPS_01F)
    p = pkg.ProteinProfile.sample(1, 2, cfg)
    q = o32.sample(1, 2, 1, EPS_01F)
    rng = np.random.default_rng(0)
    for st in (pkg.PROTEIN_N_STATE, pkg.PROTEIN_J_STATE, pkg.PROTEIN_C_STATE, pkg.PROTEIN_R_STATE, 1, 2, (1 << 14) | 1):
        for n in range(1, 6):
            for _ in range(12):
                frag = "".join("ACGT"[i] for i in rng.integers(0, 4, n))
                assert p.decode(frag, st) == q.decode(st, frag)
    assert p.decode("ACGTAC", 1)[0] == pkg.RC_EINVAL
    assert p.decode("ACN", 1)[0] == pkg.RC_EINVAL
    assert p.decode("ACG", pkg.PROTEIN_B_STATE)[0] == pkg.RC_EINVAL
    for sid in [0xC000 + i for i in range(8)] + [1, 37, 4096, (1 << 14) | 12, (2 << 14) | 255]:
        assert pkg.protein_state_name(sid) == o32.state_name(sid)
    assert pkg.protein_state_name(pkg.PROTEIN_R_STATE) == "R" and pkg.protein_state_name(2) == "M2"
    assert pkg.protein_state_is_mute(pkg.PROTEIN_S_STATE) and pkg.protein_state_is_mute((2 << 14) | 3)
    assert not pkg.protein_state_is_mute(pkg.PROTEIN_N_STATE) and not pkg.protein_state_is_mute(5)


def test_codec_reproduces_reference_codons(pkg, o32):
    """protein_codec_next over the oracle's golden path with the PRODUCT's decode (test/protein_profile.c:83-102)."""
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_UNIFORM, EPS_01F)
    p = pkg.ProteinProfile.sample(1, 2, cfg)
    rc, _, path = oracle_twin(o32, p, EPS_01F).viterbi_alt(SEQ32)
    pos, cods = 0, []
    for st, ln in path:
        if not pkg.protein_state_is_mute(st):
            rc, cod, _ = p.decode(SEQ32[pos:pos + ln], st)
            assert rc == 0
            cods.append(cod)
            pos += ln
    assert cods == GOLD["codons"]


def test_lrt(pkg):
    assert pkg.xmath_lrt(-48.0, -40.0) == 16.0
    assert pkg.xmath_lrt(np.float32(-48.927269), np.float32(-55.594276)) == np.float32(-2) * (
        np.float32(-48.927269) - np.float32(-55.594276))


def test_shard_profiles(pkg):
    rng = np.random.default_rng(3)
    sizes = np.clip(np.exp(rng.normal(np.log(130), 0.7, 5000)), 50, 2000).astype(np.uint32)
    for n in (1, 2, 4, 8):
        sh = pkg.shard_profiles(sizes, n)
        # balanced by MODELLED COST (padded width / measured class rate), not by nominal length
        cost = np.array([pkg.profile_cost(m) for m in sizes])
        loads = np.bincount(sh, weights=cost, minlength=n)
        assert sh.max() < n and loads.max() - loads.min() <= cost.max()
        assert loads.max() * n <= 1.01 * loads.sum()
    with pytest.raises(pkg.DcpError):
        pkg.shard_profiles(sizes, 0)
    # cost per node differs between kernel classes: a 3000-node profile (two-block groups) costs more per node
    assert pkg.profile_cost(3000) / 3000 > 2 * pkg.profile_cost(256) / 256
    assert pkg.profile_cost(200) == pkg.profile_cost(256)  # same padded width, same class


def test_kernel_shape_covers_every_core_size(pkg):
    """Profile length -> (warps, nodes per lane, blocks) and padded width: the padded width always holds the profile,
    one warp (or half of one) up to 256 nodes, one block up to 2048, a 2-block cluster up to 4096 (limits.h:11), never
    more than 50 % padding above 48 nodes."""
    for M in range(1, 4097):
        w, q, b = pkg.kernel_shape(M)
        pad = pkg.kernel_padded_width(M)
        assert 1 <= q <= 8 and 1 <= w <= 16 and b in (1, 2)
        assert pad >= M and pad in (w * 32 * q, 16 * q), (M, w, q, pad)
        assert (w == 1) == (M <= 256)
        assert (b == 2) == (M > 2048)
        if b == 2:
            assert w % 2 == 0 and w // 2 <= 8
        if M > 48:
            assert pad <= 1.5 * M + 16, (M, w, q, pad)
    # the chooser minimises padded width / measured rate over the class table (dcp_classes.h): the cost never
    # decreases with the profile length, and a profile never lands in a class it does not fit
    costs = [pkg.profile_cost(M) for M in range(1, 4097)]
    assert all(b >= a for a, b in zip(costs, costs[1:]))
    assert pkg.kernel_shape(256) == (1, 8, 1) and pkg.kernel_shape(512)[0] == 2 and pkg.kernel_shape(4096) == (16, 8, 2)
    assert pkg.kernel_padded_width(60) == 64 and pkg.kernel_padded_width(200) == 256  # two pairs per warp below 129 nodes
    for bad in (0, 4097):
        with pytest.raises(pkg.DcpError):
            pkg.kernel_shape(bad)


def test_no_cpu_fallback(pkg):
    """Without a CUDA device the engine refuses to exist (there is no CPU path to fall back to)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.DcpError) as e:
        pkg.Db(0)
    assert e.value.rc == pkg.RC_EFAIL


def test_plan7_inputs_build(pkg, o32):
    rng = np.random.default_rng(1)
    nl, ma, tr = plan7_profile_inputs(rng, 40)
    p = pkg.ProteinProfile.from_model(nl, ma, tr)
    assert np.isfinite(p.entry).all() and abs(np.exp(p.entry.astype(np.float64)) @ np.arange(40, 0, -1) - 1) < 1e-5


def test_h3reader_parses_hmmer3_ascii(pkg, tmp_path):
    """protein_h3reader_next over a HMMER3/f file: same tables as feeding the rounded scores directly."""
    from common import SWISSPROT_BG, write_hmm
    rng = np.random.default_rng(4)
    profs = []
    for i, M in enumerate((3, 40, 129)):
        _, ma, tr = plan7_profile_inputs(rng, M)
        profs.append(("fam%d" % i, "PF%05d.%d" % (i + 1, i + 3), ma, tr))
    path = str(tmp_path / "mini.hmm")
    seen = write_hmm(path, profs)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    got = pkg.read_hmm(path, cfg)
    assert [p.accession for p in got] == [p[1] for p in profs]
    null_lp = np.log(SWISSPROT_BG).astype(np.float32)
    for p, (ma, tr) in zip(got, seen):
        want = pkg.ProteinProfile.build(null_lp, ma, tr, cfg)
        assert p.core_size == ma.shape[0]
        assert np.array_equal(p.trans, want.trans)
        assert np.array_equal(p.match_emission, want.match_emission)
        assert np.array_equal(p.entry, want.entry) and np.array_equal(p.null_emission, want.null_emission)


def test_h3reader_rejects_garbage(pkg, tmp_path):
    bad = tmp_path / "bad.hmm"
    bad.write_text("HMMER3/f [x]\nNAME a\nLENG 2\nALPH amino\nHMM   A C\n m->m\n  1.0 2.0\n")
    with pytest.raises(pkg.DcpError) as e:
        pkg.read_hmm(str(bad))
    assert e.value.rc == pkg.RC_EPARSE
    notfmt = tmp_path / "x.hmm"
    notfmt.write_text("hello\n")
    with pytest.raises(pkg.DcpError) as e:
        pkg.read_hmm(str(notfmt))
    assert e.value.rc == pkg.RC_EPARSE
    empty = tmp_path / "empty.hmm"
    empty.write_text("")
    assert pkg.read_hmm(str(empty)) == []


def test_dcp_database_round_trip(pkg, tmp_path):
    """test/protein_db.c in spirit: write sampled profiles (seeds 1, 2), read back, same numbers."""
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    profs = [pkg.ProteinProfile.sample(1, 2, cfg, "seed1"), pkg.ProteinProfile.sample(2, 2, cfg, "seed2"),
             pkg.ProteinProfile.sample(3, 300, cfg, "PF00003.1")]
    path = str(tmp_path / "db.dcp")
    pkg.write_dcp(path, profs, cfg)
    raw = open(path, "rb").read()
    # map(2), own magic 0xC6F1: the four imm blobs of a reference file (magic 0xC6F0) are explicit arrays here
    assert raw[0] == 0x82 and raw[1:8] == b"\xa6header" and b"\xcd\xc6\xf1" in raw[:40]
    rcfg, back = pkg.read_dcp(path)
    assert rcfg.entry_dist == cfg.entry_dist and rcfg.epsilon == cfg.epsilon
    assert len(back) == 3  # EQ(db.nprofiles, 2) in the reference's test
    for a, b in zip(profs, back):
        assert a.accession == b.accession and a.core_size == b.core_size
        for f in ("match_emission", "insert_emission", "null_emission", "trans", "entry"):
            assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True)
        for w in (-2, -1, 0, a.core_size - 1):
            assert np.array_equal(a.nuclt_dist(w), b.nuclt_dist(w))
        assert a.decode("ACGT", 1) == b.decode("ACGT", 1)
    # truncated / corrupt files are parse errors, not crashes
    bad = str(tmp_path / "bad.dcp")
    open(bad, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(pkg.DcpError) as e:
        pkg.read_dcp(bad)
    assert e.value.rc == pkg.RC_EPARSE
    # a file carrying the reference's magic is refused up front, naming the reason
    open(bad, "wb").write(raw.replace(b"\xcd\xc6\xf1", b"\xcd\xc6\xf0", 1))
    with pytest.raises(pkg.DcpError) as e:
        pkg.read_dcp(bad)
    assert e.value.rc == pkg.RC_EPARSE and "imm" in str(e.value) and "reference" in str(e.value)
    open(bad, "wb").write(raw.replace(b"\xcd\xc6\xf1", b"\xcd\xc6\xf2", 1))
    with pytest.raises(pkg.DcpError) as e:
        pkg.read_dcp(bad)
    assert e.value.rc == pkg.RC_EPARSE and "magic" in str(e.value)
    with pytest.raises(pkg.DcpError):
        pkg.write_dcp(str(tmp_path / "x.dcp"), [pkg.ProteinProfile.sample(1, 2, pkg.protein_cfg(2, 0.02))], cfg)


def test_dcp_reader_is_memory_safe_on_corrupt_files(pkg, tmp_path):
    """protein_db_reader_* (src/db/protein_reader.c's role) on ~1600 truncated and corrupted databases -- structure bytes
    overwritten, lengths blown up to 2^32 - 1, container types swapped -- under AddressSanitizer and UBSan: every file
    ends in a return code, none in a memory error."""
    import os
    import random
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "deciphon-old_b200", "csrc")
    exe = str(tmp_path / "fuzz")
    build = subprocess.run(["gcc", "-std=c11", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", os.path.join(root, "include"), "-I", csrc, os.path.join(root, "tests", "dcp_reader_fuzz.c")] +
                           [os.path.join(csrc, f) for f in ("dcp_db.c", "dcp_model.c", "dcp_error.c", "dcp_shape.c")] +
                           ["-lm", "-o", exe], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizers not available: " + build.stderr[-200:])
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    good = str(tmp_path / "db.dcp")
    pkg.write_dcp(good, [pkg.ProteinProfile.sample(1, 2, cfg, "seed1"), pkg.ProteinProfile.sample(3, 40, cfg, "PF00003.1")], cfg)
    raw = open(good, "rb").read()
    rng = random.Random(11)
    files = []

    def put(b):
        path = str(tmp_path / ("f%04d" % len(files)))
        open(path, "wb").write(bytes(b))
        files.append(path)

    for c in sorted(set([0, 1, 2, 7, 8, 9, 20, 39, 40, 41, 64, 100, 200, 300, 400, 1000, len(raw) - 1] +
                        [rng.randrange(len(raw)) for _ in range(100)])):
        put(raw[:c])
    structure = [i for i in range(len(raw) - 6) if raw[i] in (0x82, 0x88, 0x90, 0x93, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9,
                                                               0xaa, 0xab, 0xac, 0xdc, 0xdd, 0xca, 0xcd, 0xce, 0xc4, 0xc5, 0xc6,
                                                               0xd9, 0xda, 0xdb)]
    for k in range(1500):
        b = bytearray(raw)
        pos = rng.randrange(min(len(raw), 400)) if k % 2 else rng.choice(structure)
        if k % 4 == 0:
            b[pos] = rng.randrange(256)
        elif k % 4 == 1:
            b[pos:pos + 4] = bytes(rng.randrange(256) for _ in range(4))
        elif k % 4 == 2:
            b[pos] = rng.choice([0xdd, 0xdc, 0xdb, 0xc6, 0xdf, 0xde, 0xcf, 0xce])
        else:
            b[pos + 1:pos + 5] = b"\xff\xff\xff\xff"
        put(b)
    put(raw)
    out = subprocess.run([exe] + files, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "AddressSanitizer" not in out.stderr and "runtime error" not in out.stderr, out.stderr[-2000:]
    assert out.stdout.startswith("files %d " % len(files))
    nerr = int(out.stdout.split()[3])
    assert 100 < nerr < len(files)  # every truncation and most structure damage is refused; the intact file is read


def test_h3reader_is_memory_safe_on_corrupt_files(pkg, tmp_path):
    """protein_h3reader_* (src/model/protein_h3reader.c + hmmer-reader's role) on ~1350 damaged HMMER3 ASCII files --
    truncations, bytes and tokens replaced ("*", "nan", 300-character tokens), lines dropped, duplicated and padded, LENG
    out of range or disagreeing with the body -- under AddressSanitizer and UBSan: a return code every time."""
    import os
    import random
    import shutil
    import subprocess
    from common import plan7_profile_inputs, write_hmm
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "deciphon-old_b200", "csrc")
    exe = str(tmp_path / "h3fuzz")
    build = subprocess.run(["gcc", "-std=c11", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", os.path.join(root, "include"), "-I", csrc, os.path.join(root, "tests", "h3reader_fuzz.c")] +
                           [os.path.join(csrc, f) for f in ("dcp_h3reader.c", "dcp_model.c", "dcp_error.c", "dcp_shape.c")] +
                           ["-lm", "-o", exe], capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizers not available: " + build.stderr[-200:])
    nrng = np.random.default_rng(3)
    models = []
    for i, M in enumerate((5, 30)):
        _, ma, tr = plan7_profile_inputs(nrng, M)
        models.append(("fam%d" % i, "PF1%04d.1" % i, ma, tr))
    good = str(tmp_path / "good.hmm")
    write_hmm(good, models)
    raw = open(good, "rb").read()
    r = random.Random(5)
    files = []

    def put(b):
        path = str(tmp_path / ("h%04d" % len(files)))
        open(path, "wb").write(bytes(b))
        files.append(path)

    for c in sorted(set([0, 1, 5, 17, 40, 100] + [r.randrange(len(raw)) for _ in range(150)])):
        put(raw[:c])
    lines = raw.split(b"\n")
    for k in range(1200):
        mode = k % 6
        if mode == 0:
            b = bytearray(raw)
            b[r.randrange(len(raw))] = r.randrange(256)
            put(b)
        elif mode == 1:
            b = bytearray(raw)
            pos = r.randrange(len(raw))
            b[pos:pos + 8] = bytes(r.choice(b" \t\n*-.0123456789eE+x") for _ in range(8))
            put(b)
        elif mode == 2:
            ls = list(lines)
            del ls[r.randrange(len(ls))]
            put(b"\n".join(ls))
        elif mode == 3:
            ls = list(lines)
            ls.insert(r.randrange(len(ls)), ls[r.randrange(len(ls))])
            put(b"\n".join(ls))
        elif mode == 4:
            ls = list(lines)
            i = r.randrange(len(ls))
            toks = ls[i].split()
            if toks:
                toks[r.randrange(len(toks))] = r.choice([b"*", b"nan", b"inf", b"-1e999", b"99999999999999999999", b"", b"x" * 300])
                ls[i] = b" ".join(toks)
            put(b"\n".join(ls))
        else:
            ls = list(lines)
            i = r.randrange(len(ls))
            ls[i] = ls[i] + b" 1.0" * r.randrange(1, 40)
            put(b"\n".join(ls))
    for leng in (b"4097", b"0", b"-5", b"29", b"31"):
        put(raw.replace(b"LENG  30", b"LENG  " + leng))
    put(raw)
    out = subprocess.run([exe] + files, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "AddressSanitizer" not in out.stderr and "runtime error" not in out.stderr, out.stderr[-2000:]
    words = out.stdout.split()
    assert int(words[1]) == len(files) and 100 < int(words[3]) < len(files) and int(words[5]) >= 2


def test_model_layer_under_sanitizers(tmp_path):
    """protein_model_* / protein_profile_* (src/model/protein_model.c, protein_profile.c) compiled with
    AddressSanitizer, LeakSanitizer and UBSan: sampled profiles of 2..70 nodes under both entry distributions, specials
    for lengths 1..10^5, decode of every fragment length in N / J / C / M / I states, a model full of zero probabilities
    (no NaN in the tables), and the error paths (setup range, add before setup, one node or transition too many,
    fragment or node out of range) -- tests/model_sanitize.c returns the number of the first check that fails."""
    import os
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "deciphon-old_b200", "csrc")
    exe = str(tmp_path / "model_san")
    build = subprocess.run(["gcc", "-std=c11", "-g", "-O1", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                            "-I", os.path.join(root, "include"), "-I", csrc, os.path.join(root, "tests", "model_sanitize.c")] +
                           [os.path.join(csrc, f) for f in ("dcp_model.c", "dcp_error.c", "dcp_shape.c")] + ["-lm", "-o", exe],
                           capture_output=True, text=True)
    if build.returncode != 0:
        pytest.skip("sanitizers not available: " + build.stderr[-200:])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stderr[-2000:])
    assert out.stdout.startswith("ok ") and "runtime error" not in out.stderr and "Sanitizer" not in out.stderr
