"""Parity of the CUDA path (through the C ABI) against the oracle on identical inputs.
Scores are compared bit for bit (same fp32 operation order), hit sets and decoded paths exactly."""
import numpy as np
import pytest

import orc
from common import (GOLD, SEQ32, SEQ1053, frameshift, oracle_twin, plan7_profile_inputs, random_seq, ref_paths,
                    sample_read)

pytestmark = pytest.mark.gpu
EPS_01F = float(np.float32(0.1))


def make_db(pkg, o, specs, eps, device=0):
    """specs: list of (seed, M, entry_dist) sampled profiles.  Returns (db, product profiles, oracle twins)."""
    db = pkg.Db(device)
    twins = []
    for seed, M, entry in specs:
        p = pkg.ProteinProfile.sample(seed, M, pkg.protein_cfg(entry, eps), accession="PF%05d" % seed)
        db.add(p)
        twins.append(oracle_twin(o, p, eps))
    db.commit()
    return db, twins


def check_scan(pkg, o, db, twins, seqs, multi_hits=True, hmmer3_compat=False, thr=10.0, flavour=0, rows=True):
    res = db.scan(seqs, multi_hits, hmmer3_compat, thr, True)
    ref = o.scan(twins, seqs, multi_hits, hmmer3_compat, thr, flavour, True)
    assert ref["rc"] == 0
    nprof = len(twins)
    assert np.array_equal(res.null_loglik, ref["null"]), np.abs(res.null_loglik - ref["null"]).max()
    assert np.array_equal(res.alt_loglik, ref["alt"]), np.abs(res.alt_loglik - ref["alt"]).max()
    assert np.array_equal(res.hit, ref["hit"])
    want = ref_paths(ref, nprof)
    assert res.nhits == int(ref["hit"].sum()) == len(want)
    order = []
    for i in range(res.nhits):
        si, pi, path = res.hit_at(i)
        order.append((si, pi))
        assert path == want[(si, pi)], (si, pi)
        assert sum(l for _, l in path) == len(seqs[si])
        if rows:
            row = res.product_row(i, scan_id=7, seq_id=100 + si)
            exp = twins[pi].product_row(7, 100 + si, "PF%05d" % 0 if False else db.profiles[pi].accession,
                                        float(ref["alt"][si, pi]), float(ref["null"][si, pi]), seqs[si], path)
            assert row == exp
    assert order == sorted(order)  # (sequence, profile) order regardless of GPU scheduling
    return res, ref


def test_reference_golden_case_on_gpu(pkg, o32):
    """test/protein_profile.c on the GPU path: loglik within the reference's fp32 tolerance, path shape, codons."""
    for entry, key in ((pkg.ENTRY_DIST_UNIFORM, "uniform"), (pkg.ENTRY_DIST_OCCUPANCY, "occupancy")):
        db, twins = make_db(pkg, o32, [(1, 2, entry)], EPS_01F)
        res = db.scan([SEQ32], True, False, -1e30, True)
        g = GOLD[key]
        assert abs(float(res.null_loglik[0, 0]) - g["null_loglik"]) <= 5e-5 * abs(g["null_loglik"])
        assert abs(float(res.alt_loglik[0, 0]) - g["alt_loglik"]) <= 5e-5 * abs(g["alt_loglik"])
        si, pi, path = res.hit_at(0)
        assert len(path) == g["alt_nsteps"] and path[0] == (pkg.PROTEIN_S_STATE, 0) and path[-1] == (pkg.PROTEIN_T_STATE, 0)
        pos, cods = 0, []
        for st, ln in path:
            if ln:
                cods.append(db.profiles[0].decode(SEQ32[pos:pos + ln], st)[1])
                pos += ln
        assert cods == GOLD["codons"]


@pytest.mark.parametrize("eps", [EPS_01F, 0.01])
def test_config1_sampled_profiles(pkg, o32, eps):
    """BASELINE config 1 re-expressed: sampler profiles x the reference's test strings (+ frameshifted)."""
    specs = [(1, 2, 1), (2, 3, 2), (3, 5, 1), (4, 17, 2), (5, 64, 2), (6, 2, 2), (7, 33, 1), (8, 64, 1)]
    db, twins = make_db(pkg, o32, specs, eps)
    rng = np.random.default_rng(1)
    seqs = [SEQ32, SEQ1053, frameshift(SEQ32, rng, 0.1), frameshift(SEQ1053, rng), random_seq(rng, 77), "A", "ACGTA"]
    for mh, h3 in ((True, False), (False, False), (True, True), (False, True)):
        check_scan(pkg, o32, db, twins, seqs, mh, h3, thr=-1e30)
    check_scan(pkg, o32, db, twins, seqs, thr=10.0)


@pytest.mark.parametrize("M", [2, 31, 32, 33, 96, 128, 129, 160, 192, 193, 200, 224, 255, 256])
def test_every_lane_layout(pkg, o32, M):
    """One profile per nodes-per-lane class boundary; short and ragged sequence lengths."""
    db, twins = make_db(pkg, o32, [(M, M, 2)], 0.01)
    rng = np.random.default_rng(M)
    seqs = [random_seq(rng, n) for n in (1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 14, 15, 16, 61, 150)]
    check_scan(pkg, o32, db, twins, seqs, thr=-1e30, rows=False)


@pytest.mark.parametrize("M", [257, 300, 320, 321, 350, 384, 385, 448, 449, 512, 513, 576, 577, 672, 673, 777, 1024, 1500, 2048, 2049, 2561, 3000, 3585, 4096])
def test_multi_warp_profiles(pkg, o32, M):
    """Profiles above 256 nodes: several warps per pair, carries exchanged through shared memory."""
    db, twins = make_db(pkg, o32, [(M, M, 2), (M + 1, 40, 2)], 0.01)
    rng = np.random.default_rng(M)
    seqs = [random_seq(rng, n) for n in (1, 2, 5, 6, 11, 64, 157)]
    check_scan(pkg, o32, db, twins, seqs, thr=-1e30, rows=False)
    check_scan(pkg, o32, db, twins, seqs[3:], multi_hits=False, hmmer3_compat=True, thr=-1e30, rows=False)


def test_long_plan7_profile_with_hit(pkg, o32):
    """A 600-node Pfam-shaped profile and a read drawn from it: long D runs across warp boundaries."""
    rng = np.random.default_rng(77)
    nl, ma, tr = plan7_profile_inputs(rng, 600)
    p = pkg.ProteinProfile.from_model(nl, ma, tr, pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01), "LONG600")
    db = pkg.Db(0)
    db.add(p)
    db.commit()
    tw = [oracle_twin(o32, p, 0.01)]
    # the read skips 40 nodes in the middle: with single-hit scoring the best path deletes them,
    # and the D run crosses the warp boundary at node 256
    full = sample_read(rng, ma, 1800, 0.0, 0.0)
    read = full[:720] + full[840:1500]
    res, ref = check_scan(pkg, o32, db, tw, [read, full[:900], random_seq(rng, 400)], multi_hits=False, thr=10.0,
                          flavour=1)
    assert res.nhits >= 2
    dels = [st & 0x3fff for st, _ in res.hit_at(0)[2] if (st >> 14) == 2]
    assert len(dels) >= 30 and min(dels) <= 256 <= max(dels) + 40
    check_scan(pkg, o32, db, tw, [read, full[:900]], multi_hits=True, thr=10.0, flavour=1)


@pytest.mark.parametrize("M,skip", [(1500, (600, 1100)), (3000, (1200, 1800)), (4096, (1700, 2700))])
def test_deletion_run_across_whole_warps_and_blocks(pkg, o32, M, skip):
    """The read matches the profile up to node skip[0] and again from skip[1]: the D chain between them runs through
    whole warps (256 nodes each) and, above 2048 nodes, from one block of the cluster into the other, so the lazy
    carry exchange needs several rounds in every row after the first matched stretch."""
    rng = np.random.default_rng(M)
    nl, ma, tr = plan7_profile_inputs(rng, M)
    p = pkg.ProteinProfile.from_model(nl, ma, tr, pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01), "DEL%d" % M)
    db = pkg.Db(0)
    db.add(p)
    db.commit()
    tw = [oracle_twin(o32, p, 0.01)]
    full = sample_read(rng, ma, 3 * M, 0.0, 0.0)
    a, b = skip
    read = full[3 * (a - 250):3 * a] + full[3 * b:3 * (b + 150)]
    for multi in (False, True):
        res, ref = check_scan(pkg, o32, db, tw, [read, full[3 * (a - 100):3 * (a + 100)]], multi_hits=multi, thr=10.0,
                              flavour=1, rows=False)
        assert res.nhits == 2


def test_plan7_profiles_with_real_hits(pkg, o32):
    """Pfam-shaped profiles, frameshifted coding reads drawn from them: real hits, long D/I stretches."""
    rng = np.random.default_rng(42)
    db = pkg.Db(0)
    twins, models = [], []
    for i, M in enumerate((200, 64, 150, 37)):
        nl, ma, tr = plan7_profile_inputs(rng, M)
        p = pkg.ProteinProfile.from_model(nl, ma, tr, pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01), "PL%d" % i)
        db.add(p)
        twins.append(oracle_twin(o32, p, 0.01))
        models.append(ma)
    db.commit()
    seqs = []
    for i in range(12):
        ma = models[i % 4]
        seqs.append(sample_read(rng, ma, int(rng.integers(300, 900)), 0.02, 0.01))
    seqs.append(random_seq(rng, 500))
    res, ref = check_scan(pkg, o32, db, twins, seqs, thr=10.0, flavour=1)
    assert 6 <= res.nhits < len(seqs) * 4  # the planted reads hit, random pairs do not


def test_multiple_hits_per_read(pkg, o32):
    """multi_hits: two copies of a domain in one read must go through J."""
    rng = np.random.default_rng(3)
    nl, ma, tr = plan7_profile_inputs(rng, 50)
    p = pkg.ProteinProfile.from_model(nl, ma, tr, pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01), "TWICE")
    db = pkg.Db(0)
    db.add(p)
    db.commit()
    tw = [oracle_twin(o32, p, 0.01)]
    read = sample_read(rng, ma, 150, 0.0, 0.0) + random_seq(rng, 40) + sample_read(rng, ma, 150, 0.01, 0.0)
    res, ref = check_scan(pkg, o32, db, tw, [read], thr=10.0)
    _, _, path = res.hit_at(0)
    assert any(st == pkg.PROTEIN_J_STATE for st, _ in path)
    res2, _ = check_scan(pkg, o32, db, tw, [read], multi_hits=False, thr=10.0)
    assert not any(st == pkg.PROTEIN_J_STATE for st, _ in res2.hit_at(0)[2])


def test_exact_ties_follow_the_documented_order(pkg, o32):
    """Degenerate profiles (every node identical, homopolymer reads) create exact score ties everywhere; the
    GPU takes the same winner as the oracle: first maximum in (transition order, source length) -- DESIGN.md 3."""
    M = 70
    null_lp = np.log(np.full(20, 0.05))
    match_lp = np.tile(np.log(np.full(20, 0.05)), (M, 1))
    t = np.log(np.array([0.90, 0.05, 0.05, 0.5, 0.5, 0.5, 0.5]))
    tr = np.tile(t, (M + 1, 1))
    tr[0][6] = -np.inf
    tr[M][2] = tr[M][6] = -np.inf
    for entry in (pkg.ENTRY_DIST_UNIFORM, pkg.ENTRY_DIST_OCCUPANCY):
        p = pkg.ProteinProfile.from_model(null_lp, match_lp, tr, pkg.protein_cfg(entry, 0.01), "TIES")
        db = pkg.Db(0)
        db.add(p)
        db.commit()
        tw = [oracle_twin(o32, p, 0.01)]
        seqs = ["A" * 90, "ACG" * 40, "AAAC" * 25, "G" * 7]
        for mh in (True, False):
            res, ref = check_scan(pkg, o32, db, tw, seqs, multi_hits=mh, thr=-1e30)
            assert res.nhits == len(seqs)


def test_scores_only_and_resident_path(pkg, o32):
    db, twins = make_db(pkg, o32, [(1, 40, 2), (2, 70, 2)], 0.01)
    rng = np.random.default_rng(9)
    seqs = [random_seq(rng, n) for n in (100, 333)]
    staged = db.stage(seqs)
    a = db.scan_resident(staged, lrt_threshold=-1e30, want_paths=False)
    b = db.scan(seqs, lrt_threshold=-1e30, want_paths=True)
    assert np.array_equal(a.alt_loglik, b.alt_loglik) and np.array_equal(a.null_loglik, b.null_loglik)
    assert a.nhits == 4 and a.hit_at(0)[2] == []
    t = b.timing
    assert t.launches >= 5 and t.alt_cells == (40 + 70) * 433 and t.score_ms > 0


def test_error_paths(pkg, o32):
    db, _ = make_db(pkg, o32, [(1, 5, 2)], 0.01)
    with pytest.raises(pkg.DcpError) as e:
        db.scan(["ACGT", ""])
    assert e.value.rc == pkg.RC_EINVAL  # protein_profile_setup(L=0) -> RC_EINVAL
    with pytest.raises(pkg.DcpError) as e:
        db.scan(["ACGN"])
    assert e.value.rc == pkg.RC_EINVAL
    with pytest.raises(pkg.DcpError):
        db.add(pkg.ProteinProfile.sample(1, 5))  # already committed
    with pytest.raises(pkg.DcpError) as e:
        pkg.ProteinProfile.sample(3, 4097, pkg.protein_cfg(2, 0.01))  # PROTEIN_MODEL_CORE_SIZE_MAX = 4096
    assert e.value.rc == pkg.RC_EINVAL
    db2 = pkg.Db(0)
    db2.add(pkg.ProteinProfile.sample(1, 5, pkg.protein_cfg(2, 0.01)))
    with pytest.raises(pkg.DcpError):
        db2.add(pkg.ProteinProfile.sample(2, 5, pkg.protein_cfg(2, 0.02)))  # one epsilon per db


def test_deterministic_across_runs(pkg, o32):
    db, _ = make_db(pkg, o32, [(s, 100 + s, 2) for s in range(1, 9)], 0.01)
    rng = np.random.default_rng(2)
    seqs = [random_seq(rng, int(n)) for n in rng.integers(50, 400, 64)]
    a = db.scan(seqs, lrt_threshold=-1e30)
    b = db.scan(seqs, lrt_threshold=-1e30)
    assert np.array_equal(a.alt_loglik, b.alt_loglik)
    assert [a.hit_at(i) for i in range(0, a.nhits, 37)] == [b.hit_at(i) for i in range(0, b.nhits, 37)]


def test_dcp_scan_driver_matches_reference_product_rows(pkg, o32, tmp_path):
    """dcp-scan (HMMER3 file + FASTA -> TSV) against the oracle's restatement of prod_fwrite."""
    import os
    import subprocess
    from common import write_hmm
    rng = np.random.default_rng(12)
    models = []
    for i, M in enumerate((60, 150, 300)):
        _, ma, tr = plan7_profile_inputs(rng, M)
        models.append(("fam%d" % i, "PF9%04d.1" % i, ma, tr))
    hmm = str(tmp_path / "db.hmm")
    seen = write_hmm(hmm, models)
    seqs = [sample_read(rng, seen[i % 3][0], int(rng.integers(200, 700)), 0.02, 0.01) for i in range(7)]
    seqs.append(random_seq(rng, 333))
    fasta = tmp_path / "reads.fasta"
    with open(fasta, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">read%d some description\n" % i)
            for a in range(0, len(s), 60):
                f.write(s[a:a + 60].lower() if i % 2 else s[a:a + 60])
                f.write("\n")
    exe = os.path.join(os.path.dirname(pkg.__file__), "dcp-scan")
    out = subprocess.run([exe, "--scan-id", "42", "--batch", "3", hmm, str(fasta)], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines(keepends=True)
    assert lines[0] == "scan_id\tseq_id\tprofile_name\tabc_name\talt_loglik\tnull_loglik\tprofile_typeid\tversion\tmatch\n"
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    profs = pkg.read_hmm(hmm, cfg)
    twins = [oracle_twin(o32, p, 0.01) for p in profs]
    ref = o32.scan(twins, seqs, True, False, 10.0, 1, True)
    paths = ref_paths(ref, 3)
    want = []
    for s in range(len(seqs)):
        for p in range(3):
            if ref["hit"][s, p]:
                want.append(twins[p].product_row(42, s + 1, profs[p].accession, float(ref["alt"][s, p]),
                                                 float(ref["null"][s, p]), seqs[s], paths[(s, p)]))
    assert len(want) >= 6
    assert lines[1:] == want
    bad = tmp_path / "bad.fasta"
    bad.write_text(">x\nACGTNNACGT\n")
    out = subprocess.run([exe, hmm, str(bad)], capture_output=True, text=True)
    assert out.returncode == 1 and "ACGT" in out.stderr


def test_scan_from_a_pressed_dcp_database(pkg, o32, tmp_path):
    """hmm_press to a .dcp file (src/server/hmm.c:120-178), then the scan loads the database from it
    (protein_db_reader_open + profile_reader_next, src/db/profile_reader.c): rows are byte-identical to the scan of the
    .hmm itself and to the oracle's product rows; the library path (read_dcp -> dcpgpu_db_add) agrees too."""
    import os
    import subprocess
    from common import write_hmm
    rng = np.random.default_rng(31)
    models = []
    for i, M in enumerate((45, 130, 260, 520)):
        _, ma, tr = plan7_profile_inputs(rng, M)
        models.append(("dom%d" % i, "PF7%04d.2" % i, ma, tr))
    hmm = str(tmp_path / "fam.hmm")
    seen = write_hmm(hmm, models)
    seqs = [sample_read(rng, seen[i % 4][0], int(rng.integers(300, 1200)), 0.02, 0.01) for i in range(10)]
    seqs.append(random_seq(rng, 500))
    fasta = tmp_path / "reads.fasta"
    fasta.write_text("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    exe = os.path.join(os.path.dirname(pkg.__file__), "dcp-scan")
    dcp = str(tmp_path / "fam.dcp")
    out = subprocess.run([exe, "--press", hmm, dcp], capture_output=True, text=True)
    assert out.returncode == 0 and "pressed 4 profiles" in out.stderr, out.stderr
    from_hmm = subprocess.run([exe, "--scan-id", "9", hmm, str(fasta)], capture_output=True, text=True)
    from_dcp = subprocess.run([exe, "--scan-id", "9", dcp, str(fasta)], capture_output=True, text=True)
    assert from_hmm.returncode == 0 and from_dcp.returncode == 0, from_dcp.stderr
    assert from_dcp.stdout == from_hmm.stdout and len(from_dcp.stdout.splitlines()) >= 9
    # the oracle's rows for the same database
    cfg, profs = pkg.read_dcp(dcp)
    assert cfg.entry_dist == pkg.ENTRY_DIST_OCCUPANCY and [p.core_size for p in profs] == [45, 130, 260, 520]
    twins = [oracle_twin(o32, p, 0.01) for p in profs]
    ref = o32.scan(twins, seqs, True, False, 10.0, 1, True)
    paths = ref_paths(ref, 4)
    want = [twins[p].product_row(9, s + 1, profs[p].accession, float(ref["alt"][s, p]), float(ref["null"][s, p]),
                                 seqs[s], paths[(s, p)])
            for s in range(len(seqs)) for p in range(4) if ref["hit"][s, p]]
    assert from_dcp.stdout.splitlines(keepends=True)[1:] == want
    # library path: profiles unpacked from the file go straight into a device database
    db = pkg.Db(0)
    for p in profs:
        db.add(p)
    db.commit()
    res = db.scan(seqs)
    assert np.array_equal(res.alt_loglik, ref["alt"]) and np.array_equal(res.hit, ref["hit"])
    # a file with the reference's magic number is refused with a parse error, not scanned
    raw = open(dcp, "rb").read()
    ref_like = str(tmp_path / "ref.dcp")
    open(ref_like, "wb").write(raw.replace(b"\xcd\xc6\xf1", b"\xcd\xc6\xf0", 1))
    out = subprocess.run([exe, ref_like, str(fasta)], capture_output=True, text=True)
    assert out.returncode == 1 and "imm" in out.stderr


def test_traceback_checkpointing_is_independent_of_its_knobs(pkg, o32, tmp_path):
    """The traceback stores a ring checkpoint every C rows, recomputes segments backwards with a band of cells and
    walks them (dcp_trace.cu).  Rows must not depend on C, on the step buffer's first size (a path that does not fit
    is retried with a larger buffer), nor on whether a segment had to be recomputed with every cell: five-row
    segments (a checkpoint before every group of rows, the walk crosses a segment boundary at almost every step), a
    segment longer than the reads, and eight-step buffers, on profiles of one half-warp, one warp and two warps, with
    a read that deletes 60 nodes (the walk leaves its band) and one with two domains (E and J)."""
    import os
    import subprocess
    from common import write_hmm
    rng = np.random.default_rng(2024)
    models = []
    for i, M in enumerate((70, 210, 400)):
        _, ma, tr = plan7_profile_inputs(rng, M)
        models.append(("tb%d" % i, "PF6%04d.1" % i, ma, tr))
    hmm = str(tmp_path / "tb.hmm")
    seen = write_hmm(hmm, models)
    seqs = [sample_read(rng, seen[i % 3][0], int(rng.integers(150, 1100)), 0.02, 0.01) for i in range(9)]
    full = sample_read(rng, seen[2][0], 1200, 0.0, 0.0)
    seqs.append(full[:450] + full[630:1100])                                   # 60 deleted nodes
    seqs.append(sample_read(rng, seen[0][0], 200, 0.0, 0.0) + random_seq(rng, 60) + sample_read(rng, seen[0][0], 200, 0.0, 0.0))
    fasta = tmp_path / "reads.fasta"
    fasta.write_text("".join(">s%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    exe = os.path.join(os.path.dirname(pkg.__file__), "dcp-scan")

    def run(*flags, **env):
        e = dict(os.environ)
        e.update({k: str(v) for k, v in env.items()})
        out = subprocess.run([exe, "--scan-id", "5", *flags, hmm, str(fasta)], capture_output=True, text=True, env=e)
        assert out.returncode == 0, out.stderr
        return out.stdout

    import re
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    profs = pkg.read_hmm(hmm, cfg)
    twins = [oracle_twin(o32, p, 0.01) for p in profs]
    for flags, multi in (((), True), (("--single-hit",), False)):
        base = run(*flags)
        ref = o32.scan(twins, seqs, multi, False, 10.0, 1, True)
        paths = ref_paths(ref, 3)
        want = [twins[p].product_row(5, s + 1, profs[p].accession, float(ref["alt"][s, p]), float(ref["null"][s, p]),
                                     seqs[s], paths[(s, p)])
                for s in range(len(seqs)) for p in range(3) if ref["hit"][s, p]]
        assert len(want) >= 10
        assert base.splitlines(keepends=True)[1:] == want
        if multi:
            assert any(",J," in r for r in want)
        else:
            assert max(len(re.findall(r",D\d+,", r)) for r in want) >= 50
        assert run(*flags, DCPGPU_TRACE_SEG=5) == base
        assert run(*flags, DCPGPU_TRACE_SEG=1000) == base
        assert run(*flags, DCPGPU_TRACE_CAP=8) == base
        assert run(*flags, DCPGPU_TRACE_SEG=10, DCPGPU_TRACE_CAP=8) == base


def test_every_kernel_class_on_short_ragged_sequences(pkg, o32):
    """One profile per row of the kernel class table (two pairs per warp, one warp, groups of 2 / 3 / 4 / 6 / 8 warps,
    two-block groups) against sequences of 1..97 nt, every pair traced: scores, paths and the (sequence, profile)
    order.  Ragged lengths put a long and a short sequence into the two halves of one warp."""
    rng = np.random.default_rng(17)
    sizes = [20, 60, 75, 90, 120, 150, 180, 250, 300, 380, 440, 500, 560, 620, 700, 800, 1000, 1100, 1300, 1700, 2000,
             2500, 3500]
    shapes = {(pkg.kernel_shape(m), pkg.kernel_padded_width(m)) for m in sizes}
    assert len(shapes) >= 20  # every row of the table is exercised
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    db = pkg.Db(0)
    twins = []
    for i, M in enumerate(sizes):
        p = pkg.ProteinProfile.build(*plan7_profile_inputs(rng, M), cfg, "CLS%02d" % i)
        db.add(p)
        twins.append(oracle_twin(o32, p, 0.01))
    db.commit()
    seqs = [random_seq(rng, n) for n in (1, 97, 7, 64, 41, 2, 33)]
    res = db.scan(seqs, lrt_threshold=-1e30)
    ref = o32.scan(twins, seqs, thr=-1e30, flavour=1)
    assert np.array_equal(res.alt_loglik, ref["alt"]) and np.array_equal(res.null_loglik, ref["null"])
    want = ref_paths(ref, len(sizes))
    assert res.nhits == len(want)
    for i in range(res.nhits):
        s, p, path = res.hit_at(i)
        assert path == want[(s, p)], (s, sizes[p])


def test_fp32_path_within_reference_tolerance_of_double_oracle(pkg, o32, o64):
    """The reference's CI runs float and double builds against the same goldens with rel. tolerance 5e-5 for float
    (test/hope_support.h:26).  The fp32 GPU scores must sit within that tolerance of the DOUBLE oracle fed the
    same fp32-rounded tables (accumulation error only), at config-2 pair shapes (M = 200, L ~ 1000)."""
    rng = np.random.default_rng(21)
    db = pkg.Db(0)
    profs, models = [], []
    for i, M in enumerate((200, 200, 120)):
        nl, ma, tr = plan7_profile_inputs(rng, M)
        p = pkg.ProteinProfile.from_model(nl, ma, tr, pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01), "T%d" % i)
        db.add(p)
        profs.append(p)
        models.append(ma)
    db.commit()
    seqs = [sample_read(rng, models[i % 3], 1000, 0.02, 0.01) for i in range(6)] + [random_seq(rng, 1000)]
    res = db.scan(seqs, lrt_threshold=10.0)
    twins64 = [oracle_twin(o64, p, 0.01) for p in profs]
    worst = 0.0
    for s in range(len(seqs)):
        for k in range(3):
            rc, nl, al = twins64[k].scores_fast(seqs[s])
            assert rc == 0
            worst = max(worst, abs(res.alt_loglik[s, k] - al) / abs(al), abs(res.null_loglik[s, k] - nl) / abs(nl))
            lrt64 = -2 * (nl - al)
            if abs(lrt64 - 10.0) > 1e-2:  # away from the threshold the hit decision agrees with double arithmetic
                assert bool(res.hit[s, k]) == (lrt64 >= 10.0)
    assert worst < 5e-5, worst


def test_zero_probability_transitions_and_emissions(pkg, o32):
    """'*' scores (probability 0) anywhere in the model: -inf arithmetic must stay NaN-free and bit-exact."""
    rng = np.random.default_rng(8)
    db = pkg.Db(0)
    twins = []
    for i, M in enumerate((40, 200, 300)):
        nl, ma, tr = plan7_profile_inputs(rng, M)
        tr = tr.copy()
        ma = ma.copy()
        for col in (1, 2, 3, 4, 5, 6):  # MI MD IM II DM DD
            idx = rng.choice(np.arange(1, M), size=max(1, M // 10), replace=False)
            tr[idx, col] = -np.inf
        ma[rng.integers(0, M, M // 5), rng.integers(0, 20, M // 5)] = -np.inf  # impossible residues
        p = pkg.ProteinProfile.from_model(nl, ma, tr, pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01), "Z%d" % i)
        assert np.isfinite(p.entry).any()
        db.add(p)
        twins.append(oracle_twin(o32, p, 0.01))
    db.commit()
    seqs = [random_seq(rng, n) for n in (7, 150, 401)] + ["ACG" * 60]
    for mh in (True, False):
        res, ref = check_scan(pkg, o32, db, twins, seqs, multi_hits=mh, thr=-1e30, rows=False)
        assert not np.isnan(res.alt_loglik).any()
