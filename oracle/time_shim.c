/* time_shim.c -- ORACLE SUPPORT (test infrastructure).
 * LD_PRELOAD shim that pins time(2) so that the `srand(time(NULL))` in front
 * of the reference's bit-flip injection (cg.cpp:79) becomes reproducible:
 *   ABFT_FAKE_TIME=12345 LD_PRELOAD=oracle/libfaketime.so cg-csr ... -x
 * Both the cpu and the cuda target then draw the same rand() sequence. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdlib.h>
#include <time.h>

time_t time(time_t *out) {
  const char *s = getenv("ABFT_FAKE_TIME");
  time_t t;
  if (s && *s) {
    t = (time_t)strtoll(s, NULL, 10);
  } else {
    time_t (*real)(time_t *) = (time_t(*)(time_t *))dlsym(RTLD_NEXT, "time");
    t = real ? real(NULL) : (time_t)0;
  }
  if (out) *out = t;
  return t;
}
