/*
 * Plain-C restatement of the correlation lookup's INDEX work (test infrastructure, see oracle/README.md).
 * Follows the reference's fp32 operation order exactly:
 *   models/utils/corr_lookup.py:115      grid = coords + flow
 *   models/utils/corr_lookup.py:127-128  c = grid / 2**level + delta        (window: first axis -> x, second -> y)
 *   models/utils/corr_lookup.py:64-65    g = c * 2. / max(W-1, 1) - 1.
 *   ATen grid_sampler (align_corners)    i = ((g + 1) / 2) * (W - 1);  i0 = floor(i)
 * Compile WITHOUT fp contraction / fast-math:  gcc -O2 -ffp-contract=off -shared -fPIC lookup_taps.c
 */
#include <math.h>
#include <stdint.h>

static float coord(float centre, int off, int size) {
  volatile float p = centre + (float)off;
  volatile float m = p * 2.0f;
  float den = (float)(size - 1 > 1 ? size - 1 : 1);
  volatile float q = m / den;
  volatile float g = q - 1.0f;
  volatile float a = g + 1.0f;
  volatile float h = a * 0.5f;
  volatile float i = h * (float)(size - 1);
  return i;
}

/* flow8: [B,H,W,2] (x,y). x0,y0: int32 [B,H,W,2r+1]; in-bounds masks mx0/mx1/my0/my1: uint8, same shape. */
void oracle_lookup_taps(const float* flow8, int B, int H, int W, int level, int radius, int32_t* x0, int32_t* y0,
                        uint8_t* mx0, uint8_t* mx1, uint8_t* my0, uint8_t* my1) {
  int hl = H, wl = W, k = 2 * radius + 1;
  float div = 1.0f;
  for (int l = 0; l < level; ++l) { hl /= 2; wl /= 2; div *= 2.0f; }
  for (long q = 0; q < (long)B * H * W; ++q) {
    int pix = (int)(q % (H * W));
    int y = pix / W, x = pix % W;
    volatile float gx = (float)x + flow8[2 * q], gy = (float)y + flow8[2 * q + 1];
    volatile float cx = gx / div, cy = gy / div;
    for (int a = 0; a < k; ++a) {
      float ix = coord(cx, a - radius, wl), iy = coord(cy, a - radius, hl);
      int xi = (int)floorf(ix), yi = (int)floorf(iy);
      x0[q * k + a] = xi;
      y0[q * k + a] = yi;
      mx0[q * k + a] = xi >= 0 && xi < wl;
      mx1[q * k + a] = xi + 1 >= 0 && xi + 1 < wl;
      my0[q * k + a] = yi >= 0 && yi < hl;
      my1[q * k + a] = yi + 1 >= 0 && yi + 1 < hl;
    }
  }
}
