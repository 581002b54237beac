/* TEST INFRASTRUCTURE. Linked with -Wl,--wrap=time into the unmodified reference programs so
 * that their `gen.seed(time(&cur_time))` (reference src/mc_eur.cpp:19, src/mc_asia.cpp:23,
 * include/common.h:192) becomes deterministic: seed = $PCF_FIXED_TIME (default 12345). */
#include <stdlib.h>
#include <time.h>
time_t __wrap_time(time_t *t) {
  const char *s = getenv("PCF_FIXED_TIME");
  time_t v = s ? (time_t)atoll(s) : (time_t)12345;
  if (t) *t = v;
  return v;
}
