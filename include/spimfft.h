/* spimfft.h -- C ABI of libspimfft.so: the Fourier-spectrum image processor of spimagine on the device.
 *
 * Replaces, for FFTProcessor.apply (spimagine/models/imageprocessor.py:82-98),
 *      res = gputools.pad_to_power2(data.astype(np.complex64), mode="wrap")
 *      res = 1./np.sqrt(res.size) * np.fft.fftshift(abs(gputools.fft(res)))
 *      res = gputools.pad_to_shape(res, dshape)            [ + np.log2(0.001 + res) ]
 * i.e. wrap-pad every axis to the next power of two (ceil(d/2) elements in front), forward 3-D FFT, magnitude,
 * fftshift, scale by 1/sqrt(#padded voxels), crop back to the volume's shape (floor(d/2) elements dropped in front).
 * Here: one pass that pads and converts the resident volume to float32, a real-to-complex cuFFT (library FFT, like
 * gputools.fft -> clFFT/reikna in the reference; half the spectrum is enough because the volume is real), and one
 * pass that reads the shifted / cropped / mirrored coefficient, takes the magnitude, scales and optionally takes the
 * logarithm.  Kept in its own library so that libspimcuda.so does not depend on cuFFT.
 *
 * Plain pointers and sizes, no exceptions; every function returns 0 or an error code (cudaError_t > 0, cufftResult
 * + 10000, SPV_E* < 0).  A plan object is not thread-safe.  There is no CPU path. */
#ifndef SPIMFFT_H
#define SPIMFFT_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SPF_API __attribute__((visibility("default")))
#else
#define SPF_API
#endif

typedef struct spf_plan spf_plan;

SPF_API int spf_create(int device, spf_plan **out);
SPF_API int spf_destroy(spf_plan *p);
/* FFTProcessor.apply.  src: C-order (nz, ny, nx) volume in host (on_device = 0) or device memory, element type
 * SPV_SRC_U8 (1), SPV_SRC_I16 (2), SPV_SRC_U16 (3) or SPV_SRC_F32 (9) of spimcuda.h.  The float32 result of the same
 * shape stays on the device (spf_result_device) and, if host_dst is not NULL, is copied into host_dst[nx*ny*nz]. */
SPF_API int spf_spectrum(spf_plan *p, const void *src, int on_device, int src_type, int nx, int ny, int nz, int take_log,
                         float *host_dst);
SPF_API int spf_result_device(spf_plan *p, float **dev);   /* valid until the next spf_spectrum / spf_destroy */
SPF_API int spf_read(spf_plan *p, float *host_dst, size_t n);  /* copy the result (n = nx*ny*nz floats) to the host */
SPF_API int spf_padded_shape(spf_plan *p, int *px, int *py, int *pz); /* the power-of-two extents of the last call */
SPF_API int spf_last_ms(spf_plan *p, float *ms);           /* device time of the last spectrum (pad + FFT + epilogue) */
SPF_API int spf_launch_count(spf_plan *p, unsigned long long *n); /* own kernels launched so far (cuFFT's not counted) */
SPF_API const char *spf_last_error(spf_plan *p);           /* p may be NULL: last create error */

#ifdef __cplusplus
}
#endif
#endif
