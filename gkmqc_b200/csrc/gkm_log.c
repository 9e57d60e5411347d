/* gkm_log.c -- level-filtered logging to fd 1 with the look of the reference's
 * logger (format "%l %d %t: %m\n", libgkm.h:27; levels gkmkern_pylib.c:118-138),
 * so that lines interleave with the Python logging of bin/gkmqc.py:222-228. */
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "gkm_internal.h"

static int g_level = GKM_LOG_ERROR;
static __thread char g_err[512];

void gkm_log_set_level(int level) { g_level = level; }
int gkm_log_get_level(void) { return g_level; }

void gkm_log(int level, const char *fmt, ...)
{
    static const char *names[] = { "TRACE", "DEBUG", "INFO", "WARN", "ERROR" };
    if (level < g_level) return;
    char buf[1400];
    time_t now = time(NULL);
    struct tm tmv;
    localtime_r(&now, &tmv);
    int n = snprintf(buf, 64, "%s ", names[level < 0 ? 0 : (level > 4 ? 4 : level)]);
    n += (int) strftime(buf + n, 40, "%Y-%m-%d %H:%M:%S: ", &tmv);
    va_list ap;
    va_start(ap, fmt);
    int m = vsnprintf(buf + n, sizeof(buf) - (size_t) n - 2, fmt, ap);
    va_end(ap);
    if (m < 0) m = 0;
    if ((size_t) m > sizeof(buf) - (size_t) n - 2) m = (int) (sizeof(buf) - (size_t) n - 2);
    n += m;
    buf[n++] = '\n';
    ssize_t r = write(1, buf, (size_t) n);
    (void) r;
}

void gkm_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    gkm_log(GKM_LOG_ERROR, "%s", g_err);
}

const char *gkmb200_last_error(void) { return g_err; }

void gkmb200_set_verbosity(int level)
{
    switch (level) {
        case 0: g_level = GKM_LOG_ERROR; break;
        case 1: g_level = GKM_LOG_WARN; break;
        case 2: g_level = GKM_LOG_INFO; break;
        case 3: g_level = GKM_LOG_DEBUG; break;
        default: g_level = GKM_LOG_TRACE; break;
    }
}
