/*
 * ref_harness_huff.c -- compiles the UNMODIFIED reference n_ary_huffman.c (where it lies, path given by
 * -DREF_HUFF_C="...") into oracle/_ref/libref_huff.so and exposes its non-static functions.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Built with -DNDEBUG: as shipped the reference aborts in its
 * first self-test (n_ary_huffman.c:916), so the as-written behaviour is only executable with
 * asserts off.  The reference narrates on stdout; the wrappers park fd 1 on /dev/null while it runs.
 * No reference source text is copied here -- the file is #included from REF_DIR at build time.
 */
#define _POSIX_C_SOURCE 200809L
#include <fcntl.h>
#include <stdio.h>
#include <unistd.h>

#define main ref_huff_main
#include REF_HUFF_C
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

/* n_ary_huffman.c:461 */
void ref_histogram(const char *text, int max_symbol_value, int *h) {
    ref_silence_begin();
    histogram(text, max_symbol_value, h);
    ref_silence_end();
}

/* n_ary_huffman.c:1161 */
void ref_huffman(int max_leaf_value, const int *freqs, int compressed_symbols, int *lengths) {
    ref_silence_begin();
    huffman(max_leaf_value, freqs, compressed_symbols, lengths);
    ref_silence_end();
}

/* n_ary_huffman.c:1382 */
void ref_convert_lengths_to_encode_table(int max_symbol_value, const int *lengths, int compressed_symbols,
                                         int *elen, unsigned int *eval) {
    ref_silence_begin();
    convert_lengths_to_encode_table(max_symbol_value, lengths, compressed_symbols, elen, eval);
    ref_silence_end();
}

/* the reference's own self-test driver (n_ary_huffman.c:2893); returns when it prints its last line */
void ref_run_tests(void) {
    ref_silence_begin();
    run_tests();
    ref_silence_end();
}

/* the reference's own block writer, static compress() (n_ary_huffman.c:1688-1815), on a buffer the caller pre-fills so that
 * what it wrote can be told from what it left alone.  With -DNDEBUG the Huffman branch (compressed_symbols > 2) runs through
 * its stubs and the function always ends in the raw pass-through block "<len>:\n\n<text>," (:1801-1814) written over the start
 * of the buffer; behind a short text the table block it had formatted before (:1705-1747) is still there. */
void ref_compress_block(int max_symbol_value, int *canonical_lengths, int compressed_symbols, int bufsize, int original_length,
                        char *original_text, char *compressed_text) {
    ref_silence_begin();
    compress(max_symbol_value, canonical_lengths, compressed_symbols, bufsize, original_length, original_text, compressed_text);
    ref_silence_end();
}
