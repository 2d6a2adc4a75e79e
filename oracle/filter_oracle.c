/*
 * filter_oracle.c -- CPU restatement of gputools.convolve_sep3, the call behind the reference's BlurProcessor /
 * BlurXYZProcessor (spimagine/models/imageprocessor.py:47-71).  TEST INFRASTRUCTURE ONLY: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product never does.
 *
 * PARITY UNPINNED.  gputools is a third-party dependency that is not vendored under /root/reference and is unpinned
 * there (setup.py:31 `"gputools"`, conda.recipe/meta.yaml:33); pyopencl is absent from this image, so neither the
 * library nor the reference can be run to produce vectors.  The algorithm below is restated from gputools' published
 * kernel file gputools/convolve/kernels/convolve_sep.cl (0.2.x), kernels conv_sep3_x / conv_sep3_y / conv_sep3_z,
 * which _convolve_sep3_gpu (gputools/convolve/convolve_sep.py) runs in that order through float32 buffers after
 * data.astype(np.float32) and h.astype(np.float32):
 *
 *     res = 0.f;
 *     h_start = (i + Nh/2 >= N) ? i + Nh/2 + 1 - N : 0;
 *     h_end   = (i - Nh/2 < 0)  ? i + Nh/2 + 1     : Nh;
 *     start   = i + Nh/2;
 *     for (ht = h_start; ht < h_end; ++ht) res += h[ht] * input[start - ht];      (along the pass axis)
 *
 * `fused` selects how `res += h * in` rounds: 1 = one fused multiply-add (what an OpenCL device compiler's default
 * contraction produces, and what libspimcuda computes), 0 = separate multiply and add.  The reference's call sites
 * pin nothing here (no test of the processors exists in the reference tree); tests/test_filters.py therefore anchors
 * the restatement on analytic known answers (impulse response = the taps, constant volume -> partial tap sums at the
 * faces, separability against a dense numpy convolution) in addition to GPU-vs-oracle equality.
 */
#include <math.h>
#include <stddef.h>

#define SFO_API __attribute__((visibility("default")))

/* one pass: axis 0 = x (fastest), 1 = y, 2 = z; in/out: nz*ny*nx floats, C order (z, y, x) */
static void conv_pass(const float *in, float *out, int nx, int ny, int nz, const float *h, int nh, int axis, int fused) {
  const int N = axis == 0 ? nx : (axis == 1 ? ny : nz);
  const size_t stride = axis == 0 ? 1 : (axis == 1 ? (size_t)nx : (size_t)nx * ny);
#pragma omp parallel for schedule(static)
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j)
      for (int i = 0; i < nx; ++i) {
        const int p = axis == 0 ? i : (axis == 1 ? j : k);
        const size_t here = ((size_t)k * ny + j) * nx + i;
        const int h_start = (p + nh / 2 >= N) ? p + nh / 2 + 1 - N : 0;
        const int h_end = (p - nh / 2 < 0) ? p + nh / 2 + 1 : nh;
        const float *line = in + here - (size_t)p * stride; /* position 0 of this line */
        const int start = p + nh / 2;
        float res = 0.f;
        if (fused) {
          for (int ht = h_start; ht < h_end; ++ht) res = fmaf(h[ht], line[(size_t)(start - ht) * stride], res);
        } else {
          for (int ht = h_start; ht < h_end; ++ht) {
            const float prod = h[ht] * line[(size_t)(start - ht) * stride];
            res = res + prod;
          }
        }
        out[here] = res;
      }
}

/* data: float32 volume (already converted like gputools' astype); tmp: scratch of the same size; the result is
 * written to `res`.  Pass order and buffer roles as in _convolve_sep3_gpu: x: data -> res, y: res -> tmp, z: tmp -> res */
SFO_API void sfo_convolve_sep3(const float *data, float *res, float *tmp, int nx, int ny, int nz, const float *hx, int nhx,
                               const float *hy, int nhy, const float *hz, int nhz, int fused) {
  conv_pass(data, res, nx, ny, nz, hx, nhx, 0, fused);
  conv_pass(res, tmp, nx, ny, nz, hy, nhy, 1, fused);
  conv_pass(tmp, res, nx, ny, nz, hz, nhz, 2, fused);
}
