/* dcp_error.c -- "log at the failure site and return the code" (include/deciphon/core/logging.h:32-72) */
#include "dcp_internal.h"
#include <string.h>

static _Thread_local char last_error[256];

void dcp_set_error(char const *msg)
{
    strncpy(last_error, msg ? msg : "", sizeof last_error - 1);
    last_error[sizeof last_error - 1] = '\0';
}

enum rc dcp_error(enum rc rc, char const *msg)
{
    dcp_set_error(msg);
    return rc;
}

char const *dcpgpu_last_error(void) { return last_error; }
