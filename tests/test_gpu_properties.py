"""Size-independent properties of the CUDA path at the shapes of BASELINE.json's configs, where the
oracle would take hours: path rescoring (bit-exact), path/length consistency, LRT rule, determinism,
agreement between device-resident and host-buffer scans, sampled cross-checks against the oracle."""
import numpy as np
import pytest

from common import oracle_twin, plan7_profile_inputs, random_seq, sample_read

pytestmark = pytest.mark.gpu
OFF = [0, 0, 4, 20, 84, 340]
ST_S, ST_N, ST_B, ST_E, ST_J, ST_C, ST_T = [0xC000 | i for i in range(1, 8)]


def code(frag):
    v = 0
    for ch in frag:
        v = v * 4 + "ACGT".index(ch)
    return OFF[len(frag)] + v


def rescore(pkg, prof, seq, path, multi_hits=True, hmmer3_compat=False):
    """Sum transition and emission scores along a path in DP order, fp32: must equal the alt loglik bit for bit."""
    f = np.float32
    rc, x = prof.setup(len(seq), multi_hits, hmmer3_compat)
    NN, CC, JJ, NB, CT, JB, RR, EJ, EC, ET, ECC, EB, EJJ = [f(v) for v in x]
    tr, ent = prof.trans, prof.entry
    emM, emI, emN = prof.match_emission, prof.insert_emission, prof.null_emission
    spec = {(ST_S, ST_N): NN, (ST_S, ST_B): NB, (ST_N, ST_N): NN, (ST_N, ST_B): NB, (ST_E, ST_T): ET, (ST_E, ST_C): ECC,
            (ST_C, ST_C): CC, (ST_C, ST_T): CT, (ST_E, ST_B): EB, (ST_E, ST_J): EJJ, (ST_J, ST_J): JJ, (ST_J, ST_B): JB}
    score, pos, prev = f(0.0), 0, None
    for st, ln in path:
        kind, k = st >> 14, st & 0x3fff
        if prev is not None:
            pk, pkk = prev >> 14, prev & 0x3fff
            if (prev, st) in spec:
                t = spec[(prev, st)]
            elif prev == ST_B and kind == 0:
                t = ent[k - 1]
            elif st == ST_E and pk in (0, 2):
                t = f(0.0)
            elif pk == 0 and kind == 0 and k == pkk + 1:
                t = tr[pkk][0]
            elif pk == 0 and kind == 1 and k == pkk:
                t = tr[pkk][1]
            elif pk == 0 and kind == 2 and k == pkk + 1:
                t = tr[pkk][2]
            elif pk == 1 and kind == 0 and k == pkk + 1:
                t = tr[pkk][3]
            elif pk == 1 and kind == 1 and k == pkk:
                t = tr[pkk][4]
            elif pk == 2 and kind == 0 and k == pkk + 1:
                t = tr[pkk][5]
            elif pk == 2 and kind == 2 and k == pkk + 1:
                t = tr[pkk][6]
            else:
                raise AssertionError("illegal transition %x -> %x" % (prev, st))
            score = f(score + f(t))
        if ln:
            c = code(seq[pos:pos + ln])
            e = emM[k - 1][c] if kind == 0 else emI[c] if kind == 1 else emN[c]
            score = f(score + f(e))
            pos += ln
        else:
            assert pkg.protein_state_is_mute(st)
        prev = st
    assert pos == len(seq)
    return score


def build_db(pkg, rng, sizes, eps=0.01):
    from concurrent.futures import ThreadPoolExecutor
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, eps)
    models = [plan7_profile_inputs(rng, M) for M in sizes]
    with ThreadPoolExecutor(16) as ex:
        profs = list(ex.map(lambda im: pkg.ProteinProfile.build(*im[1], cfg, "SYN%05d" % im[0]), enumerate(models)))
    db = pkg.Db(0)
    for p in profs:
        db.add(p)
    db.commit()
    return db, profs, models


def check_properties(pkg, db, profs, seqs, res, nsample, rng, o32=None, multi_hits=True, noracle=6, npaths=0):
    alt, null, hit = res.alt_loglik, res.null_loglik, res.hit
    lrt = np.float32(-2) * (null - alt)
    assert np.array_equal(hit.astype(bool), np.isfinite(lrt) & ~(lrt.astype(np.float64) < 10.0))  # scan_thread.c:121-123
    assert np.all(np.isfinite(alt)) and np.all(np.isfinite(null))
    # the null score depends only on (sequence, null table): all profiles share the Swiss-Prot-like background here
    assert np.all(null == null[:, :1])
    assert res.nhits == int(hit.sum())
    pick = rng.choice(res.nhits, size=min(nsample, res.nhits), replace=False) if res.nhits else []
    last = (-1, -1)
    for i in sorted(pick):
        s, p, path = res.hit_at(int(i))
        assert (s, p) > last and hit[s, p]
        last = (s, p)
        assert path[0] == (ST_S, 0) and path[-1] == (ST_T, 0)
        assert sum(l for _, l in path) == len(seqs[s])
        got = rescore(pkg, profs[p], seqs[s], path, multi_hits)
        assert got == alt[s, p], (s, p, got, alt[s, p])
    if o32 is not None:
        # pairs against the oracle itself, bit for bit: random ones (mostly misses) and as many hits
        twins = {}
        pairs = [(int(rng.integers(0, len(seqs))), int(rng.integers(0, len(profs)))) for _ in range(noracle)]
        if res.nhits:
            hs, hp = res.hits()[0], res.hits()[1]
            pairs += [(int(hs[i]), int(hp[i])) for i in rng.choice(res.nhits, size=min(noracle, res.nhits), replace=False)]
        for s, p in pairs:
            if p not in twins:
                twins[p] = oracle_twin(o32, profs[p], 0.01)
            rc, nl, al = twins[p].scores_fast(seqs[s], multi_hits, False)
            assert rc == 0 and np.float32(nl) == null[s, p] and np.float32(al) == alt[s, p], (s, p)
        # full decoded paths of a few hits against the imm-shaped generic interpreter (first-max tie order included)
        for i in (rng.choice(res.nhits, size=min(npaths, res.nhits), replace=False) if res.nhits else []):
            s, p, path = res.hit_at(int(i))
            if p not in twins:
                twins[p] = oracle_twin(o32, profs[p], 0.01)
            rc, al, want = twins[p].viterbi_alt(seqs[s], multi_hits, False)
            assert rc == 0 and np.float32(al) == alt[s, p] and path == want, (s, p)


def test_config2_full_size(pkg, o32):
    """configs[1]: 1000 profiles (M = 200) x 10 000 reads of 1 kbp = 1e7 pairs, 2e12 cells."""
    rng = np.random.default_rng(2)
    db, profs, models = build_db(pkg, rng, [200] * 1000)
    import bench
    reads = [r.decode() for r in bench.gen_reads(models, 10000, 1000, 2)]
    staged = db.stage(reads)
    res = db.scan_resident(staged)
    assert res.timing.alt_cells == 2 * 10 ** 12
    check_properties(pkg, db, profs, reads, res, 300, rng, o32, noracle=48, npaths=8)
    assert res.nhits >= 9000  # every read was drawn from one of the profiles
    res2 = db.scan(reads)  # host-buffer path, second run: identical
    assert np.array_equal(res.alt_loglik, res2.alt_loglik) and res.nhits == res2.nhits
    assert res.hit_at(res.nhits - 1) == res2.hit_at(res2.nhits - 1)


def test_config3_shape_pfam_lengths(pkg, o32):
    """configs[2] shape: Pfam-like length distribution 50..2000 (every kernel class), 1.5 kbp reads."""
    rng = np.random.default_rng(3)
    sizes = np.clip(np.exp(rng.normal(np.log(130), 0.75, 120)), 50, 2000).astype(int).tolist() + [2000, 1025, 512, 257, 50]
    db, profs, models = build_db(pkg, rng, sizes)
    reads = [sample_read(rng, models[int(rng.integers(0, len(models)))][1], 1500, 0.02, 0.01) for _ in range(96)]
    res = db.scan(reads)
    check_properties(pkg, db, profs, reads, res, 60, rng, o32, noracle=24, npaths=6)
    assert res.nhits >= 60


def test_config4_shape_long_profiles_long_contigs(pkg, o32):
    """configs[3] shape: long profiles (M ~ 3000) x 10 kbp contigs: two-block cluster kernels, large traceback.
    Scores of twelve (M ~ 3000, L = 10 kbp) pairs and one full decoded path are compared with the oracle."""
    rng = np.random.default_rng(4)
    db, profs, models = build_db(pkg, rng, [3000, 2900, 2100])
    contigs = []
    for i in range(6):
        core = sample_read(rng, models[i % 3][1], 3 * len(models[i % 3][1]) + 300, 0.01, 0.01)[:9000]
        pad = 10000 - len(core)
        contigs.append(random_seq(rng, pad // 2) + core + random_seq(rng, pad - pad // 2))
    res = db.scan(contigs)
    check_properties(pkg, db, profs, contigs, res, 18, rng, o32, noracle=6, npaths=1)
    assert res.nhits >= 6


def test_config5_shape_short_reads(pkg, o32):
    """configs[4] shape: 150 bp Illumina-like reads (substitutions 0.5 %, indels 1e-4) x Pfam-shaped profiles."""
    rng = np.random.default_rng(5)
    sizes = np.clip(np.exp(rng.normal(np.log(130), 0.7, 300)), 50, 1200).astype(int).tolist()
    db, profs, models = build_db(pkg, rng, sizes)
    reads = [sample_read(rng, models[int(rng.integers(0, len(models)))][1], 150, 1e-4, 0.005) for _ in range(4000)]
    res = db.scan(reads)
    check_properties(pkg, db, profs, reads, res, 200, rng, o32, noracle=32, npaths=12)
    assert res.timing.alt_cells == sum(sizes) * 150 * 4000
