/* Test harness (tests/test_host_model.py::test_h3reader_is_memory_safe_on_corrupt_files): reads every HMMER3 ASCII file
 * given on the command line to its end with the library's reader; built with -fsanitize=address,undefined. */
#include "dcpgpu.h"
#include <stdio.h>
int main(int argc, char **argv)
{
    int bad = 0, models = 0;
    for (int i = 1; i < argc; ++i)
    {
        FILE *fp = fopen(argv[i], "rb");
        if (!fp) return 2;
        struct protein_cfg cfg = {ENTRY_DIST_OCCUPANCY, 0.01f};
        struct protein_h3reader *r = protein_h3reader_new(cfg, fp);
        if (!r) return 3;
        enum rc rc;
        int guard = 0;
        while (!(rc = protein_h3reader_next(r)) && ++guard < 1000)
        {
            struct protein_model const *m = protein_h3reader_model(r);
            if (m && protein_h3reader_accession(r)) models++;
        }
        if (rc != RC_END) bad++;
        protein_h3reader_del(r);
        fclose(fp);
    }
    printf("files %d errors %d models %d\n", argc - 1, bad, models);
    return 0;
}
