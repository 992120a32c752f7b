/* TEST INFRASTRUCTURE ONLY -- LD_PRELOAD shim that pins time(): the reference's test programs seed their random
 * number generator with srandom(time(NULL)) (hybridtest.c:114, vtest224.c:57-58), so two runs never see the same
 * frames.  With this preloaded both the SSE2 build and the GPU build of the unchanged program draw the same data and
 * noise, and their printed counts can be compared line for line.  V224_FIXED_TIME = the value time() returns. */
#include <stdlib.h>
#include <time.h>

time_t time(time_t *t)
{
    const char *s = getenv("V224_FIXED_TIME");
    const time_t v = s ? (time_t)atoll(s) : (time_t)20140810;
    if (t) *t = v;
    return v;
}
