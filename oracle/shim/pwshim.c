/* pwshim.c - test infrastructure: LD_PRELOAD shim for the UNMODIFIED reference binary.
 * The reference finds its config file through getpwuid(getuid())->pw_dir
 * (/root/reference/src/main.cpp:266-267), not $HOME.  Rewriting the real ~/.hla-mapper is not
 * safe when several harness processes run at once (pytest-xdist), so run_reference.py gives
 * each run a private directory instead: this shim answers getpwuid() with pw_dir =
 * $HLM_REF_HOME.  Nothing in the product loads it. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pwd.h>
#include <stdlib.h>
#include <sys/types.h>

struct passwd* getpwuid(uid_t uid) {
  static struct passwd* (*real)(uid_t) = 0;
  static struct passwd copy;
  if (!real) real = (struct passwd * (*)(uid_t)) dlsym(RTLD_NEXT, "getpwuid");
  struct passwd* p = real ? real(uid) : 0;
  const char* home = getenv("HLM_REF_HOME");
  if (!p || !home || !*home) return p;
  copy = *p;
  copy.pw_dir = (char*)home;
  return &copy;
}
