"""Multi-device scans through the C ABI (dcpgpu_mdb_*): cost-weighted profile shards or sequence ranges,
one host thread per device, hits merged in (sequence, profile) order inside the library.  The merged result must
be identical -- scores, hit set, paths, product rows -- to a single-device scan and to the oracle, whatever the
device count or axis.  On a one-GPU box the shards share device 0 (an ordinal may repeat); with two or more
GPUs the same tests also run across real devices."""
import os
import subprocess

import numpy as np
import pytest

from common import oracle_twin, plan7_profile_inputs, random_seq, ref_paths, sample_read, write_hmm

pytestmark = pytest.mark.gpu


def ngpus():
    import torch
    return torch.cuda.device_count()


def build(pkg, o32, sizes, seed=21):
    rng = np.random.default_rng(seed)
    cfg = pkg.protein_cfg(pkg.ENTRY_DIST_OCCUPANCY, 0.01)
    models = [plan7_profile_inputs(rng, M) for M in sizes]
    profs = [pkg.ProteinProfile.build(*m, cfg, "MD%04d" % i) for i, m in enumerate(models)]
    twins = [oracle_twin(o32, p, 0.01) for p in profs] if o32 is not None else None
    return rng, models, profs, twins


def snapshot(res):
    seq, prof, alt, null, ns = res.hits()
    return (seq.tolist(), prof.tolist(), alt.tobytes(), null.tobytes(), ns.tolist(), res.steps().tobytes())


def device_lists():
    out = [[0], [0, 0], [0, 0, 0]]
    n = ngpus()
    if n >= 2:
        out += [[0, 1], list(range(min(n, 8)))]
    return out


@pytest.mark.parametrize("axis", [1, 2])  # AXIS_PROFILES, AXIS_SEQUENCES
def test_mdb_equals_single_device_and_oracle(pkg, o32, axis):
    sizes = [40, 300, 90, 200, 17, 520, 130, 64, 260, 700]
    rng, models, profs, twins = build(pkg, o32, sizes)
    seqs = [sample_read(rng, models[i % len(sizes)][1], int(rng.integers(150, 900)), 0.02, 0.01) for i in range(13)]
    seqs += [random_seq(rng, 77), random_seq(rng, 1)]
    db = pkg.Db(0)
    for p in profs:
        db.add(p)
    db.commit()
    single = db.scan(seqs)
    want = snapshot(single)
    ref = o32.scan(twins, seqs, True, False, 10.0, 1, True)
    assert np.array_equal(single.alt_loglik, ref["alt"]) and np.array_equal(single.hit, ref["hit"])
    paths = ref_paths(ref, len(sizes))
    assert single.nhits == len(paths) >= 10
    rows = [single.product_row(i, 3, 50 + single.hit_at(i)[0]) for i in range(single.nhits)]
    for devs in device_lists():
        m = pkg.Mdb(devs)
        for p in profs:
            m.add(p)
        m.commit(axis)
        assert m.axis == axis and m.ndevices == len(devs) and m.nprofiles == len(sizes)
        res = m.scan(seqs)
        assert snapshot(res) == want, devs
        for i in range(res.nhits):
            si, pi, path = res.hit_at(i)
            assert path == paths[(si, pi)]
        assert [res.product_row(i, 3, 50 + res.hit_at(i)[0]) for i in range(res.nhits)] == rows
        # the pair matrices are assembled lazily from the shards
        assert np.array_equal(res.alt_loglik, single.alt_loglik), devs
        assert np.array_equal(res.null_loglik, single.null_loglik)
        assert np.array_equal(res.hit, single.hit)
        parts = res.part_timings
        assert len(parts) >= 1 and all(d in devs for d, _ in parts)
        assert sum(t.alt_cells for _, t in parts) == single.timing.alt_cells
        if axis == 1:
            owners = [m.device_of(i) for i in range(len(sizes))]
            assert set(owners) <= set(devs)
        del res, m


def test_auto_axis_follows_the_cost_model(pkg, o32):
    """Pfam-like length mixes balance on the profile axis; a handful of long profiles does not."""
    _, _, profs, _ = build(pkg, None, [60, 90, 130, 150, 170, 200, 230, 250, 120, 140, 180, 110], seed=3)
    m = pkg.Mdb([0, 0])
    for p in profs:
        m.add(p)
    m.commit()
    assert m.axis == pkg.AXIS_PROFILES and m.imbalance <= 1.10
    _, _, longp, _ = build(pkg, None, [3000, 900, 700], seed=4)
    m2 = pkg.Mdb([0, 0])
    for p in longp:
        m2.add(p)
    m2.commit()
    assert m2.axis == pkg.AXIS_SEQUENCES and m2.device_of(0) == -1
    with pytest.raises(pkg.DcpError):
        m2.add(longp[0])  # committed
    with pytest.raises(pkg.DcpError):
        pkg.Mdb([0]).commit()  # empty
    with pytest.raises(pkg.DcpError):
        pkg.Mdb([999])  # no such device


def test_scan_tiled_by_device_memory_equals_one_launch(pkg, o32):
    """dcpgpu_scan splits a batch that would not fit the device's free memory into several launch sets
    (DCPGPU_SCAN_BUDGET_KB forces it here) and merges them; progress is reported per launch set."""
    sizes = [33, 150, 280]
    rng, models, profs, twins = build(pkg, o32, sizes, seed=8)
    seqs = [sample_read(rng, models[i % 3][1], int(rng.integers(100, 500)), 0.02, 0.01) for i in range(40)]
    db = pkg.Db(0)
    for p in profs:
        db.add(p)
    db.commit()
    one = db.scan(seqs)
    assert len(one.part_timings) == 0
    calls = []
    os.environ["DCPGPU_SCAN_BUDGET_KB"] = "100"  # row records are 64 B per nucleotide: a few sequences per launch set
    try:
        tiled = db.scan(seqs, progress=calls.append)
    finally:
        del os.environ["DCPGPU_SCAN_BUDGET_KB"]
    assert len(tiled.part_timings) > 3 and len(calls) == len(tiled.part_timings)
    assert sum(calls) == len(seqs) * len(sizes)
    assert snapshot(tiled) == snapshot(one)
    assert np.array_equal(tiled.alt_loglik, one.alt_loglik) and np.array_equal(tiled.hit, one.hit)
    ref = o32.scan(twins, seqs, True, False, 10.0, 1, True)
    assert np.array_equal(tiled.alt_loglik, ref["alt"])


def test_dcp_scan_devices_output_is_identical(pkg, tmp_path):
    rng = np.random.default_rng(5)
    models = []
    for i, M in enumerate((60, 150, 300, 90)):
        _, ma, tr = plan7_profile_inputs(rng, M)
        models.append(("fam%d" % i, "PF8%04d.1" % i, ma, tr))
    hmm = str(tmp_path / "db.hmm")
    seen = write_hmm(hmm, models)
    seqs = [sample_read(rng, seen[i % 4][0], int(rng.integers(200, 700)), 0.02, 0.01) for i in range(9)]
    fasta = tmp_path / "reads.fasta"
    fasta.write_text("".join(">r%d\n%s\n" % (i, s) for i, s in enumerate(seqs)))
    exe = os.path.join(os.path.dirname(pkg.__file__), "dcp-scan")
    base = subprocess.run([exe, hmm, str(fasta)], capture_output=True, text=True)
    assert base.returncode == 0, base.stderr
    assert len(base.stdout.splitlines()) >= 8
    devs = "0,1" if ngpus() >= 2 else "0,0"
    for extra in (["--devices", devs], ["--devices", devs, "--axis", "sequences"], ["--devices", "0,0,0", "--batch", "4"]):
        out = subprocess.run([exe] + extra + [hmm, str(fasta)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        assert out.stdout == base.stdout, extra
        assert "devices" in out.stderr


@pytest.mark.skipif("ngpus() < 2")
def test_two_real_devices_pfam_shape(pkg, o32):
    """Config-3 shape across real GPUs: merged N=2 (and N=all) hits and paths equal N=1; the per-device busy times
    are reported and the modelled imbalance is small."""
    rng = np.random.default_rng(9)
    sizes = np.clip(np.exp(rng.normal(np.log(130), 0.75, 160)), 50, 2000).astype(int).tolist()
    rng, models, profs, twins = build(pkg, None, sizes, seed=10)
    seqs = [sample_read(rng, models[int(rng.integers(0, len(sizes)))][1], 1500, 0.02, 0.01) for _ in range(96)]
    db = pkg.Db(0)
    for p in profs:
        db.add(p)
    db.commit()
    want = snapshot(db.scan(seqs))
    for devs in ([0, 1], list(range(min(ngpus(), 8)))):
        m = pkg.Mdb(devs)
        for p in profs:
            m.add(p)
        m.commit()
        assert m.axis == pkg.AXIS_PROFILES and m.imbalance < 1.05
        res = m.scan(seqs)
        assert snapshot(res) == want
        busy = [t.total_ms for _, t in res.part_timings]
        assert len(busy) == len(devs) and max(busy) < 1.6 * (sum(busy) / len(busy))
        del res, m
