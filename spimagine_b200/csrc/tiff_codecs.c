/* Strip decoders for compressed TIFF stacks (host side, plain C, no CUDA): the frame reader in front of the
 * upload path.  The reference gets these from `tifffile` (spimagine/utils/imgutils.py:18-23,
 * spimagine/models/data_model.py:178-218); the algorithms are the ones TIFF 6.0 publishes: section 13 (LZW, with
 * the "early change" of the code width every TIFF writer uses) and section 9 (PackBits).
 *
 * Every function returns 0 and the number of bytes produced in *written, or a negative code:
 *   -1 bad argument, -2 the stream is damaged, -3 the stream holds more than `cap` bytes (the first `cap` bytes
 *   are still delivered: writers may pad the last strip to RowsPerStrip).
 * They are reentrant and are called through ctypes with the GIL released, so the reader thread decodes one time
 * point while the renderer works on another.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "spimtiff.h"

#define LZW_CLEAR 256
#define LZW_EOI 257
#define LZW_FIRST 258
#define LZW_MAX 4096

int spt_version(void) { return 100; }

int spt_lzw_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *written) {
    if (!src || !dst || !written) return -1;
    /* Every table string is the previous code's string plus one byte, and that is exactly what lies in the output
     * from where the previous string was written: an entry is (position in the output, length), and a code is
     * decoded by copying forwards inside dst -- no prefix chains, no per-byte table walk. */
    size_t where[LZW_MAX];
    uint32_t length[LZW_MAX];
    size_t out = 0, at = 0, prev_at = 0;
    uint32_t prev_len = 0;
    uint64_t acc = 0;
    int have = 0, width = 9, next = LZW_FIRST, prev = -1;
    *written = 0;
    for (;;) {
        while (have < width && at < n) {
            acc = (acc << 8) | src[at++];
            have += 8;
        }
        if (have < width) break; /* no end-of-information code: libtiff accepts that as well */
        int code = (int)((acc >> (have - width)) & ((1u << width) - 1));
        have -= width;
        if (code == LZW_EOI) break;
        if (code == LZW_CLEAR) {
            width = 9;
            next = LZW_FIRST;
            prev = -1;
            continue;
        }
        size_t from;
        uint32_t len;
        if (code < 256) {
            from = 0;
            len = 1;
        } else if (prev < 0) {
            return -2; /* a table code right behind a clear code */
        } else if (code < next) {
            from = where[code];
            len = length[code];
        } else if (code == next && next < LZW_MAX) {
            from = prev_at; /* the string being defined: previous string + its own first byte */
            len = prev_len + 1;
        } else {
            return -2;
        }
        if (prev >= 0 && next < LZW_MAX) {
            where[next] = prev_at;
            length[next] = prev_len + 1;
            next++;
            if (next + 1 >= (1 << width) && width < 12) width++;
        }
        uint32_t take = len;
        int overflow = out + len > cap;
        if (overflow) take = (uint32_t)(cap - out);
        if (code < 256) {
            if (take) dst[out] = (uint8_t)code;
        } else if (take > 32 && from + take <= out) {
            memcpy(dst + out, dst + from, take);
        } else {
            for (uint32_t k = 0; k < take; k++) dst[out + k] = dst[from + k]; /* overlaps its own output */
        }
        if (overflow) {
            *written = cap;
            return -3;
        }
        prev_at = out;
        prev_len = len;
        out += len;
        prev = code;
    }
    *written = out;
    return 0;
}

int spt_packbits_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *written) {
    if (!src || !dst || !written) return -1;
    size_t at = 0, out = 0;
    *written = 0;
    while (at < n) {
        int h = (int8_t)src[at++];
        if (h == -128) continue;
        if (h >= 0) {
            size_t run = (size_t)h + 1;
            if (at + run > n) return -2;
            if (out + run > cap) {
                memcpy(dst + out, src + at, cap - out);
                *written = cap;
                return -3;
            }
            memcpy(dst + out, src + at, run);
            at += run;
            out += run;
        } else {
            size_t run = (size_t)(1 - h);
            if (at >= n) return -2;
            if (out + run > cap) {
                memset(dst + out, src[at], cap - out);
                *written = cap;
                return -3;
            }
            memset(dst + out, src[at++], run);
            out += run;
        }
    }
    *written = out;
    return 0;
}

/* Predictor = 2 (TIFF 6.0 section 14): every sample of a row was stored as the difference to its left neighbour,
 * in wrap-around arithmetic of the sample's own width.  `swap` != 0 when the file's byte order is not the
 * machine's: the samples are then summed in the file's order and left in the file's order. */
#define UNDIFF(T, SWAP)                                           \
    for (size_t r = 0; r < rows; r++) {                           \
        T *p = (T *)(data + r * width * sizeof(T));               \
        T run = 0;                                                \
        for (size_t i = 0; i < width; i++) {                      \
            T v = p[i];                                           \
            if (swap) v = SWAP(v);                                \
            run = (T)(run + v);                                   \
            p[i] = swap ? SWAP(run) : run;                        \
        }                                                         \
    }
#define SAME(v) (v)

int spt_undo_differencing(uint8_t *data, size_t rows, size_t width, int bytes_per_sample, int swap) {
    if (!data && rows != 0 && width != 0) return -1;
    switch (bytes_per_sample) {
    case 1: { UNDIFF(uint8_t, SAME) return 0; }
    case 2: { UNDIFF(uint16_t, __builtin_bswap16) return 0; }
    case 4: { UNDIFF(uint32_t, __builtin_bswap32) return 0; }
    case 8: { UNDIFF(uint64_t, __builtin_bswap64) return 0; }
    default: return -1;
    }
}
