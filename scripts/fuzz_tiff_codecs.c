/* AddressSanitizer / UBSan fuzz of the strip decoders (random, low-entropy and clear-code-led streams, random caps):
 *   gcc -O1 -g -fsanitize=address,undefined -Iinclude scripts/fuzz_tiff_codecs.c spimagine_b200/csrc/tiff_codecs.c -o /tmp/fuzz && /tmp/fuzz
 * last run: 200000 streams, no report ("lzw ok 77951 damaged 120059 clipped 1990"). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "spimtiff.h"
int main(void) {
    srand(1);
    long ok = 0, bad = 0, over = 0;
    for (int it = 0; it < 200000; it++) {
        size_t n = rand() % 400, cap = rand() % 500;
        uint8_t *src = malloc(n ? n : 1), *dst = malloc(cap ? cap : 1);
        int mode = rand() % 3;
        for (size_t i = 0; i < n; i++) src[i] = mode == 0 ? rand() : (mode == 1 ? rand() % 4 : (rand() % 8 ? 0x80 | (rand() & 0x7f) : rand()));
        if (n && (rand() & 1)) { src[0] = 0x80; }  /* starts with a clear code (9 bits: 1 0000 0000) */
        size_t w = 0;
        int rc = spt_lzw_decode(src, n, dst, cap, &w);
        if (w > cap) { printf("overrun\n"); return 1; }
        rc == 0 ? ok++ : rc == -2 ? bad++ : over++;
        rc = spt_packbits_decode(src, n, dst, cap, &w);
        if (w > cap) { printf("overrun\n"); return 1; }
        if (cap >= 8 && n >= 8) spt_undo_differencing(dst, 1, cap / 8, 1 << (rand() % 4), rand() & 1);
        free(src); free(dst);
    }
    printf("lzw ok %ld damaged %ld clipped %ld\n", ok, bad, over);
    return 0;
}
