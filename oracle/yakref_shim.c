/*
 * yakref_shim.c — TEST INFRASTRUCTURE ONLY.
 * Thin exports over the UNMODIFIED reference C sources (compiled where they lie
 * under /root/reference/yak by oracle/Makefile, never copied): the yak hash
 * functions (yak/yak-priv.h:10-38) and table restore/get (yak/htab.c:80,213).
 * Used to pin the oracle's hash + lookup restatement (tests/test_oracle_yak.py)
 * and to generate tests/golden/yak_kat.json.
 */
#include <stdint.h>
#include "yak-priv.h"

uint64_t yakref_hash64(uint64_t key, uint64_t mask) { return yak_hash64(key, mask); }
uint64_t yakref_hash64_64(uint64_t key) { return yak_hash64_64(key); }
uint64_t yakref_hash_long(uint64_t *x) { return yak_hash_long(x); }
void *yakref_restore(const char *fn) { return yak_ch_restore(fn); }
int yakref_get(const void *h, uint64_t x) { return yak_ch_get((const yak_ch_t *)h, x); }
void yakref_destroy(void *h) { yak_ch_destroy((yak_ch_t *)h); }
int yakref_k(const void *h) { return ((const yak_ch_t *)h)->k; }
