/* spimtiff.h -- C ABI of libspimtiff.so: strip decoders for compressed TIFF stacks on the frame reader thread.
 *
 * Replaces, for read3dTiff / TiffData.load (spimagine/utils/imgutils.py:18-23, spimagine/models/data_model.py:178-218),
 * what `tifffile.imread` decodes: LZW (Compression = 5, TIFF 6.0 section 13), PackBits (32773, section 9) and the
 * horizontal-differencing predictor (Predictor = 2, section 14).  Deflate strips (8 / 32946) go through zlib from the
 * host language.  Host code only: the decoded time point lands in the page-locked buffer the upload path reads
 * (spimagine_b200/frames.py FrameSource), so no CUDA is involved and the library builds with gcc alone.
 *
 * Plain pointers and sizes; every function returns 0 or a negative code: -1 bad argument, -2 damaged stream,
 * -3 the stream holds more than `cap` bytes (the first `cap` bytes are delivered).  All functions are reentrant. */
#ifndef SPIMTIFF_H
#define SPIMTIFF_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPT_API __attribute__((visibility("default")))
#else
#define SPT_API
#endif

SPT_API int spt_version(void);

/* one strip of `n` compressed bytes -> at most `cap` bytes at dst; *written = bytes produced */
SPT_API int spt_lzw_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *written);
SPT_API int spt_packbits_decode(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, size_t *written);

/* in place: rows x width samples of 1 / 2 / 4 / 8 bytes become running sums along each row (wrap-around);
 * swap != 0: the samples are in the other byte order than the machine's and stay so */
SPT_API int spt_undo_differencing(uint8_t *data, size_t rows, size_t width, int bytes_per_sample, int swap);

#ifdef __cplusplus
}
#endif
#endif
