/* Test harness (tests/test_host_model.py::test_dcp_reader_is_memory_safe_on_corrupt_files): opens every file given on the
 * command line with the library's .dcp reader and reads all profiles; built with -fsanitize=address,undefined. */
#include "dcpgpu.h"
#include <stdio.h>
int main(int argc, char **argv)
{
    int bad = 0;
    for (int i = 1; i < argc; ++i)
    {
        FILE *fp = fopen(argv[i], "rb");
        if (!fp) return 2;
        struct protein_db_reader *r = NULL;
        enum rc rc = protein_db_reader_open(&r, fp);
        if (!rc)
        {
            for (;;)
            {
                struct protein_profile *p = NULL;
                rc = protein_db_reader_next(r, &p);
                if (rc || !p) break;
                protein_profile_del(p);
            }
            protein_db_reader_close(r);
            if (rc == RC_END) rc = RC_OK; /* read to the end */
        }
        if (rc) bad++;
        fclose(fp);
    }
    printf("files %d errors %d\n", argc - 1, bad);
    return 0;
}
