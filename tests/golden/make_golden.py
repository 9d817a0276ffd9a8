"""Regenerates the golden fixtures from the reference's own test sources (run in the dev container only;
/root/reference does not exist on the GPU box).  Writes golden.json next to this script."""
import json
import os
import re

REF = "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))

src = open(os.path.join(REF, "test/protein_h3reader.c")).read()
seq1053 = "".join(re.findall(r'^\s+"([ACGT]+)"', src, re.M))
assert len(seq1053) == 1053

pp = open(os.path.join(REF, "test/protein_profile.c")).read()
seq32 = re.search(r'char const str\[\] = "([ACGT]+)"', pp).group(1)
closes = [float(x) for x in re.findall(r"CLOSE\(prod\.loglik, (-[0-9.]+)\)", pp)]
codons = re.findall(r'IMM_CODON\(nuclt, "([ACGT]{3})"\)', pp)
nsteps = [int(x) for x in re.findall(r"EQ\(imm_path_nsteps\(&prod\.path\), (\d+)\)", pp)]
tol = re.search(r"rel_tol = ([0-9.e-]+)", open(os.path.join(REF, "test/hope_support.h")).read())

gold = {
    "source": "test/protein_profile.c, test/protein_h3reader.c, test/hope_support.h",
    "seq32": seq32,
    "seq1053": seq1053,
    "uniform": {"null_loglik": closes[0], "alt_loglik": closes[1], "null_nsteps": nsteps[0], "alt_nsteps": nsteps[1]},
    "occupancy": {"null_loglik": closes[2], "alt_loglik": closes[3], "null_nsteps": nsteps[2], "alt_nsteps": nsteps[3]},
    "codons": codons[:10],
    "codons_occupancy": codons[10:20],
    "pf02545_alt_loglik_not_runnable_offline": -1430.9281381240353,
    "epsilon_literal": "0.1f",
    "seed": 1,
    "core_size": 2,
}
json.dump(gold, open(os.path.join(here, "golden.json"), "w"), indent=1)
print(gold["uniform"], gold["occupancy"], gold["codons"], len(seq1053))
