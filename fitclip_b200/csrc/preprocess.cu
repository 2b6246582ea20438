// Evaluation pre-processing on the GPU: what the reference does per frame on CPU DataLoader
// workers -- aligner/encoder/clip_video_text_encoder.py:124-133:
//   ConvertBHWCtoBCHW -> ConvertImageDtype(float) [x / 255] -> Resize(size, BICUBIC) [shorter side -> size, no
//   antialias for tensors in the pinned torchvision 0.12] -> CenterCrop(size) -> Normalize(mean, std)
// fused into one pass: uint8 (F, H, W, 3) -> fp32 / bf16 (F, 3, size, size), only the cropped window is computed.
// The arithmetic restates ATen's upsample_bicubic2d (align_corners = false, A = -0.75, border-replicated taps).
#include <math.h>

#include "kernels.cuh"

namespace fc {

namespace {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  c[0] = cubic2(t + 1.f, A);
  c[1] = cubic1(t, A);
  const float u = 1.f - t;
  c[2] = cubic1(u, A);
  c[3] = cubic2(u + 1.f, A);
}

// Bilinear resize (Resize's default for the SLIP wrapper, slip_video_text_encoder.py:78-87) runs through the same four taps
// with weights (0, 1 - t, t, 0): ATen's upsample_bilinear2d clamps the source coordinate at 0 and the second tap at the
// last row / column, which is what the border-replicated taps give (0.3 a + 0.7 a = a up to one rounding).
__device__ __forceinline__ void linear_coeffs(float t, float (&c)[4]) {
  c[0] = 0.f;
  c[1] = 1.f - t;
  c[2] = t;
  c[3] = 0.f;
}

struct PreParams {
  int bilinear;    // 0: bicubic (CLIP's transform), 1: bilinear (SLIP's)
  int H, W;        // source frame
  int RH, RW;      // resized frame (shorter side == size)
  int top, left;   // crop offset inside the resized frame
  int size;        // output side
  float scale_h, scale_w;
  float mean[3], stdv[3];
  int P, G, ldp;   // PATCH mode: patch size, patches per side, row stride of the patch matrix (elements)
};

template <typename TOut>
__device__ __forceinline__ void store_px(TOut* p, float v);
template <>
__device__ __forceinline__ void store_px<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_px<bf16>(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// One thread = one output pixel (all three channels); consecutive threads walk an output row, so the three channel
// planes are written with fully coalesced stores.  The four horizontal taps of a source row are 12 consecutive bytes
// (4 pixels x RGB): they are fetched as four aligned 32-bit words and realigned with funnel shifts -- 16 loads per
// output pixel instead of 48 byte loads (the kernel is LSU / ALU-bound, not HBM-bound: neighbouring threads re-read the
// same lines from L1).  Pixels whose taps are clamped at the left / right border take the byte path.
// PATCH = true writes the result straight into the patch-embedding GEMM's A operand instead of an NCHW image: pixel
// (c, oy, ox) of frame f lands in row f*G*G + (oy/P)*G + ox/P, column c*P*P + (oy%P)*P + ox%P of the bf16 patch matrix
// (the layout im2col_kernel produces, elementwise.cu) -- the normalised frame never exists in memory.  The P
// consecutive pixels of a patch row are P consecutive bf16 (full 32-byte sectors for P = 16).
template <typename TOut, bool PATCH>
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ in, TOut* __restrict__ out,
                                                         int64_t frames, const PreParams p) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const int64_t f = blockIdx.z;
  if (ox >= p.size) return;
  // source coordinates of this output pixel in the resized (un-cropped) frame
  const float ry = p.scale_h * (static_cast<float>(oy + p.top) + 0.5f) - 0.5f;
  const float rx = p.scale_w * (static_cast<float>(ox + p.left) + 0.5f) - 0.5f;
  const float fy = floorf(ry), fx = floorf(rx);
  const int iy = static_cast<int>(fy), ix = static_cast<int>(fx);
  float cy[4], cx[4];
  if (p.bilinear) {
    linear_coeffs(ry - fy, cy);
    linear_coeffs(rx - fx, cx);
  } else {
    cubic_coeffs(ry - fy, cy);
    cubic_coeffs(rx - fx, cx);
  }
  const float inv255 = 1.f / 255.f;  // x * (1/255) is within 1 ulp of torch's x / 255 (tolerance: tests/test_gpu_preprocess.py)
  const uint8_t* src = in + f * static_cast<int64_t>(p.H) * p.W * 3;
  const bool interior = ix >= 1 && ix + 2 <= p.W - 2;  // no clamping, and the aligned reads stay inside the row
  float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int y = min(max(iy - 1 + i, 0), p.H - 1);
    const uint8_t* row = src + static_cast<int64_t>(y) * p.W * 3;
    float r[3] = {0.f, 0.f, 0.f};
    if (interior) {
      const uintptr_t b = reinterpret_cast<uintptr_t>(row + (ix - 1) * 3);
      const uint32_t* a = reinterpret_cast<const uint32_t*>(b & ~static_cast<uintptr_t>(3));
      const uint32_t sh = static_cast<uint32_t>(b & 3) * 8;
      const uint32_t w0 = __ldg(a), w1 = __ldg(a + 1), w2 = __ldg(a + 2), w3 = __ldg(a + 3);
      const uint32_t v[3] = {__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh)};
#pragma unroll
      for (int k = 0; k < 12; ++k) {  // byte k = pixel k / 3, channel k % 3
        // uint8 -> fp32 WITHOUT the I2F instruction (it runs on the quarter-rate XU pipe, and 48 of them per output pixel
        // were the kernel's bound): one PRMT drops the byte into the mantissa of 2^23, one FADD removes the 2^23 -- exact
        const float t = __uint_as_float(__byte_perm(v[k >> 2], 0x4B000000u, 0x7440u | (k & 3))) - 8388608.f;
        r[k % 3] = fmaf(t * inv255, cx[k / 3], r[k % 3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int xs = min(max(ix - 1 + j, 0), p.W - 1) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] = fmaf(static_cast<float>(__ldg(row + xs + c)) * inv255, cx[j], r[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) acc[c] = fmaf(r[c], cy[i], acc[c]);
  }
  if (PATCH) {
    const int py = oy / p.P, px = ox / p.P;
    TOut* dst = out + ((f * p.G + py) * p.G + px) * static_cast<int64_t>(p.ldp) + (oy - py * p.P) * p.P + (ox - px * p.P);
#pragma unroll
    for (int c = 0; c < 3; ++c) store_px<TOut>(dst + c * p.P * p.P, (acc[c] - p.mean[c]) / p.stdv[c]);
  } else {
    const int64_t plane = static_cast<int64_t>(p.size) * p.size;
    TOut* dst = out + f * 3 * plane + static_cast<int64_t>(oy) * p.size + ox;
#pragma unroll
    for (int c = 0; c < 3; ++c) store_px<TOut>(dst + c * plane, (acc[c] - p.mean[c]) / p.stdv[c]);
  }
}

}  // namespace

// patch > 0: `out` is the bf16 patch matrix [F * (size/patch)^2, ldp] (see preprocess_kernel<.., true>)
static int preprocess_impl(const uint8_t* frames, int64_t F, int H, int W, int size, const float* mean, const float* stdv,
                           void* out, int out_dtype, int patch, int ldp, int interpolation, cudaStream_t s) {
  FC_REQUIRE(frames && out && mean && stdv, "preprocess: null pointer");
  FC_REQUIRE(interpolation == 0 || interpolation == 1, "preprocess: interpolation must be 0 (bicubic) or 1 (bilinear), got %d",
             interpolation);
  FC_REQUIRE(F >= 0 && H > 0 && W > 0 && size > 0, "preprocess: bad shape F=%lld H=%d W=%d size=%d",
             static_cast<long long>(F), H, W, size);
  FC_REQUIRE(F <= 65535, "preprocess: at most 65535 frames per call (got %lld)", static_cast<long long>(F));
  if (F == 0) return FC_OK;
  PreParams p;
  p.bilinear = interpolation;
  p.H = H;
  p.W = W;
  p.size = size;
  // torchvision _compute_resized_output_size: the shorter side becomes `size`, the longer int(size * long / short)
  if (W <= H) {
    p.RW = size;
    p.RH = static_cast<int>(static_cast<double>(size) * H / W);
  } else {
    p.RH = size;
    p.RW = static_cast<int>(static_cast<double>(size) * W / H);
  }
  // torchvision center_crop: int(round((h - crop) / 2.0)) with Python's round-half-to-even
  p.top = static_cast<int>(nearbyint((p.RH - size) / 2.0));
  p.left = static_cast<int>(nearbyint((p.RW - size) / 2.0));
  p.scale_h = static_cast<float>(H) / static_cast<float>(p.RH);
  p.scale_w = static_cast<float>(W) / static_cast<float>(p.RW);
  for (int c = 0; c < 3; ++c) {
    p.mean[c] = mean[c];
    p.stdv[c] = stdv[c];
    FC_REQUIRE(stdv[c] != 0.f, "preprocess: std[%d] is zero", c);
  }
  p.P = patch;
  p.G = patch > 0 ? size / patch : 0;
  p.ldp = ldp;
  ProfScope prof(s, PROF_OTHER, patch > 0 ? 2 : 1, F, H, W, 0.0,
                 static_cast<double>(F) * (3.0 * H * W + 3.0 * size * size * (out_dtype == FC_DTYPE_F32 ? 4 : 2)));
  dim3 grid((size + 255) / 256, size, static_cast<unsigned>(F));
  if (patch > 0) {
    FC_REQUIRE(out_dtype == FC_DTYPE_BF16 && size % patch == 0 && ldp >= 3 * patch * patch,
               "preprocess (patch mode): bf16 output, size %% patch == 0 and ldp >= 3 * patch^2 required");
    preprocess_kernel<bf16, true><<<grid, 256, 0, s>>>(frames, static_cast<bf16*>(out), F, p);
  } else if (out_dtype == FC_DTYPE_F32) {
    preprocess_kernel<float, false><<<grid, 256, 0, s>>>(frames, static_cast<float*>(out), F, p);
  } else if (out_dtype == FC_DTYPE_BF16) {
    preprocess_kernel<bf16, false><<<grid, 256, 0, s>>>(frames, static_cast<bf16*>(out), F, p);
  } else {
    FC_REQUIRE(false, "preprocess: output dtype must be fp32 (0) or bf16 (1), got %d", out_dtype);
  }
  FC_CHECK_LAUNCH();
  return FC_OK;
}

int preprocess_frames(const uint8_t* frames, int64_t F, int H, int W, int size, const float* mean, const float* stdv,
                      void* out, int out_dtype, int interpolation, cudaStream_t s) {
  return preprocess_impl(frames, F, H, W, size, mean, stdv, out, out_dtype, 0, 0, interpolation, s);
}

int preprocess_to_patches(const uint8_t* frames, int64_t F, int H, int W, int size, int patch, const float* mean,
                          const float* stdv, bf16* patches, int ldp, int interpolation, cudaStream_t s) {
  return preprocess_impl(frames, F, H, W, size, mean, stdv, patches, FC_DTYPE_BF16, patch, ldp, interpolation, s);
}

}  // namespace fc
