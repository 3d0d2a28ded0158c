/*
 * ref_harness_nybble.c -- compiles the UNMODIFIED reference nybble_compression.c (path given by
 * -DREF_NYBBLE_C="...") into oracle/_ref/libref_nybble.so and exposes its functions.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  The reference prints per symbol
 * (nybble_compression.c:784-787, :952-981); fd 1 is parked on /dev/null while it runs.
 */
#define _POSIX_C_SOURCE 200809L
#include <fcntl.h>
#include <stdio.h>
#include <unistd.h>

#define main ref_nybble_main
#include REF_NYBBLE_C
#undef main

static int g_saved_fd = -1;
static int g_depth = 0;

void ref_silence_begin(void) {
    if (g_depth++ > 0) return;
    fflush(stdout);
    g_saved_fd = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    if (nul >= 0) { dup2(nul, 1); close(nul); }
}

void ref_silence_end(void) {
    if (--g_depth > 0) return;
    fflush(stdout);
    if (g_saved_fd >= 0) { dup2(g_saved_fd, 1); close(g_saved_fd); g_saved_fd = -1; }
}

/* nybble_compression.c:1091 */
void ref_write_nybble(int nybble, char *dest, int nybble_offset) { write_nybble(nybble, dest, nybble_offset != 0); }

/* nybble_compression.c:887 */
void ref_compress_bytestring(const char *src, char *dst, int modify) {
    ref_silence_begin();
    compress_bytestring(src, dst, modify != 0);
    ref_silence_end();
}

/* nybble_compression.c:734 */
void ref_decompress_bytestring(const char *src, char *dst, int modify) {
    ref_silence_begin();
    decompress_bytestring(src, dst, modify != 0);
    ref_silence_end();
}

/* the reference's own self-test main (nybble_compression.c:1139) */
int ref_nybble_selftest(void) {
    ref_silence_begin();
    int rc = ref_nybble_main();
    ref_silence_end();
    return rc;
}
