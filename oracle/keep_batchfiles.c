/* TEST INFRASTRUCTURE ONLY.  LD_PRELOAD shim for tests/golden/make_golden_cli.py.
 *
 * The reference deletes its intermediate batchfiles when a run ends (IS_DELETE_CACHE_BATCHFILE is a
 * `static const bool = true`, src/basetype_caller.h:24, used at basetype_caller.cpp:220-224,256-258) through
 * std::filesystem::remove -> remove(3).  To capture the rows the UNMODIFIED binary feeds to `_basevar_caller`,
 * this shim turns remove()/unlink()/rmdir() of anything under a "/cache_" directory into a no-op.
 *   gcc -O2 -fPIC -shared -o oracle/_ref/libkeepbf.so oracle/keep_batchfiles.c -ldl
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

static int kept(const char* p) { return p && strstr(p, "/cache_") != NULL; }

int remove(const char* p) {
    static int (*real)(const char*) = NULL;
    if (!real) real = (int (*)(const char*))dlsym(RTLD_NEXT, "remove");
    return kept(p) ? 0 : real(p);
}
int unlink(const char* p) {
    static int (*real)(const char*) = NULL;
    if (!real) real = (int (*)(const char*))dlsym(RTLD_NEXT, "unlink");
    return kept(p) ? 0 : real(p);
}
int rmdir(const char* p) {
    static int (*real)(const char*) = NULL;
    if (!real) real = (int (*)(const char*))dlsym(RTLD_NEXT, "rmdir");
    return kept(p) ? 0 : real(p);
}
