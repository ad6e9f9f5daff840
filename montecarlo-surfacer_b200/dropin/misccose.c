/* dropin/misccose.c — the three helpers the reference includes from a file it does not ship
 * (SMC.h:20).  Signatures follow their call sites: currentTime() SMC.c:122 / main.c:64,
 * new_strtof(argv, NULL, 10) main.c:18, make_directory(name) main.c:55,59. */
#ifndef SMCB_DROPIN_MISCCOSE_C
#define SMCB_DROPIN_MISCCOSE_C
#include <stdlib.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <time.h>

/* {hour, minute} of the local time; the buffer is static, as the callers never free it */
static inline int *currentTime(void)
{
    static int hour_minute[2];
    const time_t now = time(NULL);
    struct tm parts;
    localtime_r(&now, &parts);
    hour_minute[0] = parts.tm_hour;
    hour_minute[1] = parts.tm_min;
    return hour_minute;
}

/* decimal string -> double; the third argument of the only call site is ignored */
static inline double new_strtof(const char *text, char **endp, int unused_base)
{
    (void)unused_base;
    return strtod(text, endp);
}

/* mkdir that tolerates an existing directory */
static inline int make_directory(const char *path)
{
    return mkdir(path, 0775);
}
#endif
