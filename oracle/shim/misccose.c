/* oracle/shim/misccose.c — TEST INFRASTRUCTURE (oracle build only).
 *
 * The reference does `#include "misccose.c"` (SMC.h:20) but does not ship the
 * file.  These are the three symbols it needs (SMC.c:122, main.c:18,55,59,64),
 * written from their call sites.  Nothing here is on the hot path.
 */
#include <stdlib.h>
#include <time.h>
#include <sys/stat.h>
#include <sys/types.h>

static int *currentTime(void)
{
    static int hm[2];
    time_t t = time(NULL);
    struct tm lt;
    localtime_r(&t, &lt);
    hm[0] = lt.tm_hour;
    hm[1] = lt.tm_min;
    return hm;
}

/* main.c:18 calls it as new_strtof(argv[4], NULL, 10) and stores a double */
static double new_strtof(const char *s, char **end, int base_unused)
{
    (void)base_unused;
    return strtod(s, end);
}

static int make_directory(const char *name)
{
    return mkdir(name, 0777);
}
