"""Shared helpers for the parity tests."""
import numpy as np

import json
import os

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))
SEQ32 = GOLD["seq32"]      # test/protein_profile.c:27
SEQ1053 = GOLD["seq1053"]  # test/protein_h3reader.c:6-24

AMINO = "ACDEFGHIKLMNPQRSTVWY"
_TCAG = "TCAG"
_AAS = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"
CODONS_OF = {}
for _i, _a in enumerate(_TCAG):
    for _j, _b in enumerate(_TCAG):
        for _k, _c in enumerate(_TCAG):
            CODONS_OF.setdefault(_AAS[_i * 16 + _j * 4 + _k], []).append(_a + _b + _c)


def frameshift(seq, rng, indel=0.02, sub=0.01):
    out = []
    for ch in seq:
        r = rng.random()
        if r < indel / 2:
            continue
        if r < indel:
            out.append(ch)
            out.append("ACGT"[rng.integers(0, 4)])
            continue
        if r < indel + sub:
            out.append("ACGT"[rng.integers(0, 4)])
            continue
        out.append(ch)
    return "".join(out) or "A"


def random_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def plan7_profile_inputs(rng, M, sharp=4.0):
    """Pfam-shaped synthetic model inputs: peaked match distributions, Plan7-like transitions."""
    bg = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092,
                   0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364,
                   0.0540687, 0.0673417, 0.0114135, 0.0304133])
    null_lp = np.log(bg / bg.sum())
    g = rng.gamma(1.0 / sharp, 1.0, size=(M, 20)) * bg + 1e-6
    match_lp = np.log(g / g.sum(1, keepdims=True))
    n = M + 1
    mm = rng.uniform(0.85, 0.97, n)
    mi = (1 - mm) * rng.uniform(0.3, 0.7, n)
    md = 1 - mm - mi
    im = rng.uniform(0.4, 0.8, n)
    dm = rng.uniform(0.3, 0.8, n)
    with np.errstate(divide="ignore"):
        tr = np.log(np.stack([mm, mi, md, im, 1 - im, dm, 1 - dm], axis=1))
        tr[0, 6] = -np.inf  # node 0 has no delete state
        tr[0, 5] = 0.0
        tr[M] = np.log(np.array([mm[M] + md[M], mi[M], 0.0, im[M], 1 - im[M], 1.0, 0.0]))  # last node: no M->D, D->D
    return null_lp, match_lp, tr


def sample_read(rng, match_lp, L, indel=0.02, sub=0.01):
    """A read carrying a frameshifted codon path drawn from the profile, padded with random nt."""
    M = match_lp.shape[0]
    cod = []
    for k in range(M):
        p = np.exp(match_lp[k])
        aa = AMINO[rng.choice(20, p=p / p.sum())]
        cs = CODONS_OF[aa]
        cod.append(cs[rng.integers(0, len(cs))])
    core = frameshift("".join(cod), rng, indel, sub)
    if len(core) >= L:
        s = rng.integers(0, len(core) - L + 1)
        return core[s:s + L]
    pad = L - len(core)
    left = rng.integers(0, pad + 1)
    return random_seq(rng, left) + core + random_seq(rng, pad - left)


def oracle_twin(o, p, eps):
    """Oracle profile holding bit-identical DP-level numbers of a product profile."""
    M = p.core_size
    return o.import_tables(M, float(np.float32(eps)), p.match_emission, p.insert_emission, p.null_emission,
                           p.trans, p.entry, p.nuclt_dist(-2), p.nuclt_dist(-1),
                           np.stack([p.nuclt_dist(k) for k in range(M)]))


def ref_paths(ref, nprof):
    """dict (seq, prof) -> [(state, len)] from orc.Oracle.scan output."""
    out = {}
    off = ref["path_off"]
    nseq = ref["hit"].shape[0]
    for s in range(nseq):
        for p in range(nprof):
            i = s * nprof + p
            if off[i + 1] > off[i]:
                out[(s, p)] = list(zip(ref["step_state"][off[i]:off[i + 1]].tolist(),
                                       ref["step_len"][off[i]:off[i + 1]].tolist()))
    return out


SWISSPROT_BG = np.array([0.0787945, 0.0151600, 0.0535222, 0.0668298, 0.0397062, 0.0695071, 0.0229198, 0.0590092,
                         0.0594422, 0.0963728, 0.0237718, 0.0414386, 0.0482904, 0.0395639, 0.0540978, 0.0683364,
                         0.0540687, 0.0673417, 0.0114135, 0.0304133])  # HMMER3 background (protein_h3reader.c:79-103)


def _score(lp):
    return "*" if not np.isfinite(lp) else "%.5f" % (-lp)


def write_hmm(path, profiles):
    """Write HMMER3/f ASCII profiles.  profiles: [(name, acc, match_lp[M,20], trans[M+1,7])].
    Returns, per profile, the (match, trans) arrays as a reader will see them (5-decimal rounding)."""
    seen = []
    with open(path, "w") as f:
        for name, acc, ma, tr in profiles:
            M = ma.shape[0]
            f.write("HMMER3/f [3.1b2 | February 2015]\nNAME  %s\nACC   %s\nDESC  synthetic\nLENG  %d\nALPH  amino\n"
                    "RF    no\nMM    no\nCONS  yes\nCS    no\nMAP   yes\nNSEQ  10\nEFFN  1.5\nCKSUM 1\n"
                    "STATS LOCAL MSV       -9.0000  0.70000\n" % (name, acc, M))
            f.write("HMM          " + "        ".join(AMINO) + "   \n")
            f.write("            m->m     m->i     m->d     i->m     i->i     d->m     d->d\n")
            f.write("  COMPO   " + "  ".join("%.5f" % 2.9 for _ in range(20)) + "\n")
            f.write("          " + "  ".join("%.5f" % 2.9 for _ in range(20)) + "\n")
            f.write("          " + "  ".join(_score(x) for x in tr[0]) + "\n")
            for k in range(M):
                cons = AMINO[int(np.argmax(ma[k]))].lower()
                f.write("%7d   " % (k + 1) + "  ".join(_score(x) for x in ma[k]) + " %6d %s - - -\n" % (k + 1, cons))
                f.write("          " + "  ".join("%.5f" % 2.9 for _ in range(20)) + "\n")
                f.write("          " + "  ".join(_score(x) for x in tr[k + 1]) + "\n")
            f.write("//\n")
            rnd = lambda a: np.array([[-float(_score(x)) if np.isfinite(x) else -np.inf for x in row] for row in a])
            seen.append((rnd(ma), rnd(tr)))
    return seen
