/* Test harness (tests/test_host_model.py::test_model_layer_under_sanitizers): the host model layer -- sample, build with
 * zero probabilities, setup for a range of lengths, decode every fragment length in every kind of state, the error
 * paths -- built with -fsanitize=address,undefined. */
#include "dcpgpu.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(void)
{
    int checks = 0;
    for (int entry = 1; entry <= 2; ++entry)
        for (unsigned M = 2; M <= 70; M += (M < 8 ? 1 : 23))
        {
            struct protein_cfg cfg = {(enum entry_dist)entry, entry == 1 ? 0.1f : 0.01f};
            struct protein_profile *p = protein_profile_new("PF00001.1", cfg);
            if (!p) return 2;
            if (protein_profile_sample(p, 7u + M, M)) return 3;
            float sp[13];
            if (protein_profile_setup(p, 0, true, false, sp) != RC_EINVAL) return 4;
            for (unsigned L = 1; L <= 100000; L *= 7)
                for (int f = 0; f < 4; ++f)
                    if (protein_profile_setup(p, L, f & 1, f & 2, sp)) return 5;
            char const *frags[] = {"A", "CG", "TGA", "ACGT", "GATTA"};
            unsigned states[] = {PROTEIN_N_STATE, PROTEIN_J_STATE, PROTEIN_C_STATE, PROTEIN_MATCH_STATE | 1,
                                 PROTEIN_MATCH_STATE | M, PROTEIN_INSERT_STATE | 1};
            for (unsigned s = 0; s < sizeof states / sizeof states[0]; ++s)
                for (unsigned l = 1; l <= 5; ++l)
                {
                    char codon[3], amino = 0;
                    enum rc rc = protein_profile_decode(p, frags[l - 1], l, states[s], codon, &amino);
                    if (rc && rc != RC_EINVAL) return 6;
                    checks++;
                }
            /* out-of-range requests are errors, not reads */
            char codon[3];
            if (!protein_profile_decode(p, "ACGTAC", 6, PROTEIN_N_STATE, codon, NULL)) return 7;
            if (!protein_profile_decode(p, "A", 1, PROTEIN_MATCH_STATE | (M + 1), codon, NULL)) return 8;
            if (!protein_profile_decode(p, "A", 1, PROTEIN_B_STATE, codon, NULL)) return 9;
            if (!protein_profile_decode(p, "N", 1, PROTEIN_N_STATE, codon, NULL)) return 10;
            double nd[129];
            if (protein_profile_nuclt_dist(p, -2, nd) || protein_profile_nuclt_dist(p, -1, nd) ||
                protein_profile_nuclt_dist(p, (int)M - 1, nd))
                return 11;
            if (!protein_profile_nuclt_dist(p, (int)M, nd) || !protein_profile_nuclt_dist(p, -3, nd)) return 12;
            protein_profile_del(p);
        }
    /* a model with zero probabilities (-inf scores) everywhere they may occur */
    {
        unsigned M = 6;
        float null_lp[20], match[6][20], trans[7][7];
        for (int a = 0; a < 20; ++a) null_lp[a] = logf(0.05f);
        for (unsigned k = 0; k < M; ++k)
            for (int a = 0; a < 20; ++a) match[k][a] = a == (int)k ? 0.0f : -INFINITY;
        for (unsigned k = 0; k <= M; ++k)
            for (int t = 0; t < 7; ++t) trans[k][t] = (t == 0 || t == 3 || t == 5) ? 0.0f : -INFINITY;
        struct protein_cfg cfg = {ENTRY_DIST_OCCUPANCY, 0.01f};
        struct protein_profile *p = protein_profile_new("ZERO", cfg);
        if (protein_profile_build(p, M, null_lp, &match[0][0], &trans[0][0], NULL)) return 13;
        float const *e = protein_profile_match_emission(p);
        for (unsigned i = 0; i < M * 1364; ++i)
            if (isnan(e[i])) return 14;
        protein_profile_del(p);
        /* error paths of the model: setup range, add before setup, overflow */
        struct protein_model *m = protein_model_new(cfg, null_lp);
        if (protein_model_setup(m, 0) != RC_EINVAL || protein_model_setup(m, 4097) != RC_EINVAL) return 15;
        protein_model_del(m);
        m = protein_model_new(cfg, null_lp);
        if (!protein_model_add_node(m, match[0], 'A')) return 16;
        if (protein_model_setup(m, 1)) return 17;
        struct protein_trans t0 = {0};
        if (protein_model_add_trans(m, t0) || protein_model_add_node(m, match[0], 'A') || protein_model_add_trans(m, t0))
            return 18;
        if (!protein_model_add_node(m, match[0], 'A') || !protein_model_add_trans(m, t0)) return 19;
        protein_model_del(m);
    }
    printf("ok %d\n", checks);
    return 0;
}
