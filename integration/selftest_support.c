/*
 * selftest_support.c - only for the stand-alone self-test build of the shim (integration/Makefile). Inside the reference
 * tree these three symbols come from libsrsran_phy (lib/src/phy/utils/debug.c, phy_logger.c, lib/src/phy/fec/cbsegm.c).
 */
#include <stdarg.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include "srsran_b200.h"

bool is_handler_registered(void) { return true; }
void srsran_phy_log_print(int log_level, const char* format, ...)
{
  (void)log_level;
  va_list ap;
  va_start(ap, format);
  vfprintf(stderr, format, ap);
  fputc('\n', stderr);
  va_end(ap);
}
int srsran_cbsegm_cbindex(uint32_t long_cb)
{
  int i = srsb200_cbindex(long_cb);
  return (i >= 0 && (uint32_t)srsb200_cbsize((uint32_t)i) == long_cb) ? i : (i >= 0 ? i : -1);
}
