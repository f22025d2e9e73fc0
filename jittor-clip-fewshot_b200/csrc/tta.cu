// TTA view generator: from decoded uint8 images to the [n_views, 3, S, S] uint8 crop batch the image tower
// consumes -- the step immediately upstream of the hot path.
//
// The reference builds 1 centre view + N random crops per test image with PIL on the CPU, 8 DataLoader
// workers, 512 PIL operations per image (test.py:1547-1560):
//   centre view   Resize(256, BICUBIC) on the short side + CenterCrop(224)      jclip/clip.py:102-135
//   crops         RandomResizedCrop(224, BILINEAR) + RandomHorizontalFlip       test.py:1898-1903, ood.py:1084-1089
// Both are "take a box of the source, resample it with a separable filter, keep an S x S window, maybe
// mirror it".  The resample is Pillow's ImagingResample (src/libImaging/Resample.c), reproduced here BIT FOR
// BIT on uint8: double-precision filter weights normalised per output pixel, rounded to 22-bit fixed point,
// a horizontal pass into a uint8 intermediate, then a vertical pass (the checker is Pillow itself,
// tests/test_gpu_tta.py; numpy restatement in oracle/crops.py).
//
//   resample_h_kernel   one CTA = 128 intermediate rows x S columns of one view; weights for the S columns are
//                       computed once per CTA in fp64 (explicitly rounded ops: no FMA contraction) into smem
//   resample_v_kernel   one CTA = all S output rows x S columns of one view; reads the planar uint8 intermediate
//                       coalesced, writes the planar [3, S, S] uint8 view (mirrored if asked) -- or, fused with the
//                       tower's front end (OUT_PATCH_*), the view's rows of the conv1 patch matrix directly: ToTensor's
//                       1/255, tfm_clip's (x - mean) / std (test.py:1301) and the im2col of jclip/model.py:105-108 applied
//                       to the finished uint8 pixel, so the uint8 view tensor and the im2col pass over it disappear
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include "kernels.h"
#include "ptx.cuh"

namespace jcb {

namespace {

constexpr int PREC = 32 - 8 - 2;  // Pillow PRECISION_BITS
constexpr int RBH = 128;          // intermediate rows per CTA in the horizontal pass (amortises the fp64 weight set-up)
constexpr int RBV = 256;          // output rows per CTA in the vertical pass: the whole view, one weight set per thread

struct AxisDev {          // one axis of one view, as Pillow's precompute_coeffs sees it
  double scale, ss, support;   // in/out ratio, 1 / filterscale, filter support * filterscale
  int in_size, ksize, off;     // source extent along the axis (the crop), taps per output, window offset
};

struct ViewDev {
  long long src_off;      // byte offset of the source image (HWC uint8) in the packed buffer
  long long tmp_off;      // byte offset of this view's intermediate (planar [3][n_rows][S]) in the scratch
  int src_w;              // source image width (row pitch = 3 * src_w bytes)
  int top, left;          // crop origin
  int row0, n_rows;       // first intermediate row (relative to the crop) and how many the vertical pass needs
  int filter, flip;
  AxisDev h, v;
};

__device__ __forceinline__ double filt(int kind, double x) {
  if (x < 0.0) x = -x;
  if (kind == 0) return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;                       // bilinear_filter
  if (x < 1.0)                                                                    // bicubic_filter, a = -0.5
    return __dadd_rn(__dmul_rn(__dmul_rn(__dsub_rn(__dmul_rn(1.5, x), 2.5), x), x), 1.0);
  if (x < 2.0)
    return __dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dsub_rn(x, 5.0), x), 8.0), x), 4.0), -0.5);
  return 0.0;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for ONE output index X of an axis.
// Writes the taps to k[0], k[kstride], ... (ksize of them, zero padded), returns (xmin, n) through the references.
__device__ __forceinline__ void axis_coeffs(const AxisDev& a, int kind, int X, int* k, int kstride, int& xmin, int& n) {
  const double center = __dmul_rn(__dadd_rn(static_cast<double>(X), 0.5), a.scale);   // in0 = 0
  int lo = __double2int_rz(__dadd_rn(__dsub_rn(center, a.support), 0.5));
  if (lo < 0) lo = 0;
  int hi = __double2int_rz(__dadd_rn(__dadd_rn(center, a.support), 0.5));
  if (hi > a.in_size) hi = a.in_size;
  const int cnt = hi - lo;
  double ww = 0.0;
  for (int x = 0; x < cnt; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + lo), center), 0.5), a.ss);
    ww = __dadd_rn(ww, filt(kind, arg));
  }
  for (int x = 0; x < cnt; ++x) {
    const double arg = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + lo), center), 0.5), a.ss);
    double w = filt(kind, arg);
    if (ww != 0.0) w = __ddiv_rn(w, ww);
    const double f = __dmul_rn(w, static_cast<double>(1 << PREC));
    k[x * kstride] = w < 0.0 ? __double2int_rz(__dadd_rn(-0.5, f)) : __double2int_rz(__dadd_rn(0.5, f));
  }
  for (int x = cnt; x < a.ksize; ++x) k[x * kstride] = 0;
  xmin = lo;
  n = cnt;
}

// where the vertical pass leaves a view
enum TtaOut : int { OUT_U8 = 0, OUT_PATCH_BF16 = 1, OUT_PATCH_F16 = 2 };
struct PatchEmit {        // OUT_PATCH_*: patch geometry and the per-channel affine map of the pixel (rowwise.cu im2col_kernel)
  int P, Gp;              // patch edge, patches per row (S / P)
  float a[3], b[3];       // pixel -> a[c] * u8 + b[c]
};

__device__ __forceinline__ uint32_t clip8(int v) {
  v >>= PREC;
  return static_cast<uint32_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// byte B (compile-time) of a little-endian word array
template <int B, int N>
__device__ __forceinline__ int byte_of(const uint32_t (&a)[N]) {
  return static_cast<int>(__byte_perm(a[B >> 2], 0u, 0x4440u + (B & 3)));   // PRMT: byte B & 3, zero extended
}

// ---------------------------------------------------------------------------------- horizontal pass
// One thread = one intermediate pixel (row r, column xx), all three channels.  The taps of a pixel are 3 * n
// CONTIGUOUS bytes of the interleaved source row, at an arbitrary alignment: they are fetched as aligned 32-bit words
// and realigned with funnel shifts, so a 7-tap pixel costs 6 loads instead of 21 single-byte loads (the byte-load
// version was bound by the LSU, profiles/r01e_tta_full.md).  Weights are zero padded to KW, bytes past the last tap
// are multiplied by zero (and never loaded past the last word that holds a tap).
template <int KW>
__device__ __forceinline__ void h_rows_fast(const uint8_t* __restrict__ row0p, long long pitch, const int* __restrict__ kk,
                                            const int* __restrict__ xmin, const int* __restrict__ cnt, int S, int r_begin,
                                            int r_end, uint8_t* __restrict__ dst, long long plane) {
  constexpr int NA = (3 * KW + 3) / 4;   // aligned words that hold the 3 * KW tap bytes
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // a lane keeps one output column (its KW weights and window live in registers) and walks down the rows of the strip
  for (int xx = lane; xx < S; xx += 32) {
    int wt[KW];
#pragma unroll
    for (int x = 0; x < KW; ++x) wt[x] = kk[x * S + xx];
    const uint8_t* colp = row0p + 3LL * xmin[xx];
    const int nb = 3 * cnt[xx] + 3;
    uint8_t* d = dst + xx;
    for (int r = r_begin + warp; r < r_end; r += nwarps) {
      const uintptr_t addr = reinterpret_cast<uintptr_t>(colp + r * pitch);
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(addr & ~static_cast<uintptr_t>(3));
      const int mis = static_cast<int>(addr & 3);
      const int nw = (mis + nb) >> 2;
      uint32_t w[NA + 1];
#pragma unroll
      for (int j = 0; j <= NA; ++j) w[j] = j < nw ? __ldg(wp + j) : 0u;
      uint32_t a[NA];
#pragma unroll
      for (int j = 0; j < NA; ++j) a[j] = __funnelshift_r(w[j], w[j + 1], 8 * mis);
      int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
      // static unroll over the taps: the byte positions must be compile-time constants (one PRMT per byte)
      auto tap = [&](auto X) {
        constexpr int x = decltype(X)::value;
        a0 += byte_of<3 * x + 0>(a) * wt[x];
        a1 += byte_of<3 * x + 1>(a) * wt[x];
        a2 += byte_of<3 * x + 2>(a) * wt[x];
      };
      tap(std::integral_constant<int, 0>{});
      tap(std::integral_constant<int, 1>{});
      tap(std::integral_constant<int, 2>{});
      if constexpr (KW > 3) { tap(std::integral_constant<int, 3>{}); tap(std::integral_constant<int, 4>{}); }
      if constexpr (KW > 5) { tap(std::integral_constant<int, 5>{}); tap(std::integral_constant<int, 6>{}); }
      if constexpr (KW > 7) { tap(std::integral_constant<int, 7>{}); tap(std::integral_constant<int, 8>{}); }
      uint8_t* dr = d + static_cast<long long>(r) * S;
      dr[0] = static_cast<uint8_t>(clip8(a0));
      dr[plane] = static_cast<uint8_t>(clip8(a1));
      dr[2 * plane] = static_cast<uint8_t>(clip8(a2));
    }
  }
}

// any tap count: byte loads, run-time loop (large source images: scale > 4)
__device__ __forceinline__ void h_rows_generic(const uint8_t* __restrict__ row0p, long long pitch, const int* __restrict__ kk,
                                               const int* __restrict__ xmin, const int* __restrict__ cnt, int S,
                                               int r_begin, int r_end, uint8_t* __restrict__ dst, long long plane) {
  for (int p = threadIdx.x; p < (r_end - r_begin) * S; p += blockDim.x) {
    const int r = r_begin + p / S, xx = p % S;
    const uint8_t* s = row0p + r * pitch + 3LL * xmin[xx];
    const int n = cnt[xx];
    int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < n; ++x) {
      const int wt = kk[x * S + xx];
      a0 += static_cast<int>(__ldg(s + 3 * x + 0)) * wt;
      a1 += static_cast<int>(__ldg(s + 3 * x + 1)) * wt;
      a2 += static_cast<int>(__ldg(s + 3 * x + 2)) * wt;
    }
    uint8_t* d = dst + static_cast<long long>(r) * S + xx;
    d[0] = static_cast<uint8_t>(clip8(a0));
    d[plane] = static_cast<uint8_t>(clip8(a1));
    d[2 * plane] = static_cast<uint8_t>(clip8(a2));
  }
}

__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, const ViewDev* __restrict__ views, int S, int kmax,
                  uint8_t* __restrict__ tmp, int words_ok) {
  extern __shared__ __align__(16) int rs_smem[];
  const ViewDev v = views[blockIdx.y];
  const int r_begin = blockIdx.x * RBH;
  if (r_begin >= v.n_rows) return;
  int* kk = rs_smem;                 // [kmax][S]: tap x of column xx at kk[x * S + xx] (a warp reads consecutive words)
  int* xmin = kk + S * kmax;         // [S]
  int* cnt = xmin + S;               // [S]
  for (int xx = threadIdx.x; xx < S; xx += blockDim.x) {
    int lo, n;
    axis_coeffs(v.h, v.filter, xx + v.h.off, kk + xx, S, lo, n);
    xmin[xx] = lo;
    cnt[xx] = n;
  }
  __syncthreads();
  const int r_end = min(r_begin + RBH, v.n_rows);
  const long long pitch = 3LL * v.src_w;
  const long long plane = static_cast<long long>(v.n_rows) * S;
  const uint8_t* row0p = src + v.src_off + (v.top + v.row0) * pitch + 3LL * v.left;
  uint8_t* dst = tmp + v.tmp_off;
  switch (words_ok ? v.h.ksize : 0) {   // ksize = 2 * ceil(support) + 1: odd, uniform over the CTA
    case 3: h_rows_fast<3>(row0p, pitch, kk, xmin, cnt, S, r_begin, r_end, dst, plane); break;
    case 5: h_rows_fast<5>(row0p, pitch, kk, xmin, cnt, S, r_begin, r_end, dst, plane); break;
    case 7: h_rows_fast<7>(row0p, pitch, kk, xmin, cnt, S, r_begin, r_end, dst, plane); break;
    case 9: h_rows_fast<9>(row0p, pitch, kk, xmin, cnt, S, r_begin, r_end, dst, plane); break;
    default: h_rows_generic(row0p, pitch, kk, xmin, cnt, S, r_begin, r_end, dst, plane);
  }
}

// ---------------------------------------------------------------------------------- vertical pass
// One thread = four adjacent output pixels of one channel: every tap is one aligned 32-bit load of the planar
// intermediate (a warp reads 128 contiguous bytes), the four results leave as one 32-bit store (byte-reversed and
// mirrored in x for a flipped view).  Needs S % 4 == 0 and 4-byte aligned buffers; otherwise the byte version runs.
template <int KW, int OUT>
__device__ __forceinline__ void v_rows_fast(const uint8_t* __restrict__ t, long long plane, const int* __restrict__ kk,
                                            int kmax, const int* __restrict__ ymin, const int* __restrict__ cnt, int S,
                                            int y_begin, int rows, int flip, uint8_t* __restrict__ o, const PatchEmit& pe) {
  const int G = S >> 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const long long plane_w = plane >> 2;          // plane = n_rows * S, a multiple of 4
  // a warp keeps one output row (its KW weights live in registers) and sweeps the 3 * G words of the three channels
  for (int yy = warp; yy < rows; yy += nwarps) {
    const int n = cnt[yy];
    int wt[KW];
#pragma unroll
    for (int y = 0; y < KW; ++y) wt[y] = kk[yy * kmax + y];
    const uint32_t* base = reinterpret_cast<const uint32_t*>(t + static_cast<long long>(ymin[yy]) * S);
    for (int e = lane; e < 3 * G; e += 32) {
      const int c = (e >= G) + (e >= 2 * G);
      const int g = e - c * G;
      const uint32_t* s = base + c * plane_w + g;
      int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
      for (int y = 0; y < KW; ++y) {
        if (y < n) {                 // rows past the window may lie outside the intermediate
          const uint32_t w = __ldg(s + y * G);
          a0 += static_cast<int>(__byte_perm(w, 0u, 0x4440u)) * wt[y];
          a1 += static_cast<int>(__byte_perm(w, 0u, 0x4441u)) * wt[y];
          a2 += static_cast<int>(__byte_perm(w, 0u, 0x4442u)) * wt[y];
          a3 += static_cast<int>(w >> 24) * wt[y];
        }
      }
      uint32_t r4 = clip8(a0) | (clip8(a1) << 8) | (clip8(a2) << 16) | (clip8(a3) << 24);
      int go = g;
      if (flip) {                       // RandomHorizontalFlip acts on the finished S x S crop
        r4 = __byte_perm(r4, 0u, 0x0123);
        go = G - 1 - g;
      }
      if (OUT == OUT_U8) {
        reinterpret_cast<uint32_t*>(o + (static_cast<long long>(c) * S + y_begin + yy) * S)[go] = r4;
      } else {
        // the four pixels (x = 4 go .. 4 go + 3 of row y) are four consecutive columns of one patch row: one 8-byte store
        const int y = y_begin + yy, x0 = 4 * go;
        const long long prow = static_cast<long long>(y / pe.P) * pe.Gp + x0 / pe.P;
        const int col = (c * pe.P + y % pe.P) * pe.P + x0 % pe.P;
        const float pa = pe.a[c], pb = pe.b[c];
        const float f0 = fmaf(static_cast<float>(r4 & 0xFFu), pa, pb), f1 = fmaf(static_cast<float>((r4 >> 8) & 0xFFu), pa, pb);
        const float f2 = fmaf(static_cast<float>((r4 >> 16) & 0xFFu), pa, pb), f3 = fmaf(static_cast<float>(r4 >> 24), pa, pb);
        uint2 h;
        h.x = pack_h2<OUT == OUT_PATCH_F16>(f0, f1);
        h.y = pack_h2<OUT == OUT_PATCH_F16>(f2, f3);
        *reinterpret_cast<uint2*>(o + (prow * (3LL * pe.P * pe.P) + col) * 2) = h;
      }
    }
  }
}

template <int OUT>
__device__ __forceinline__ void v_rows_generic(const uint8_t* __restrict__ t, long long plane, const int* __restrict__ kk,
                                               int kmax, const int* __restrict__ ymin, const int* __restrict__ cnt, int S,
                                               int y_begin, int rows, int flip, uint8_t* __restrict__ o, const PatchEmit& pe) {
  for (int p = threadIdx.x; p < rows * S; p += blockDim.x) {
    const int yy = p / S, xx = p % S;
    const int* k = kk + yy * kmax;
    const int n = cnt[yy];
    const uint8_t* s = t + static_cast<long long>(ymin[yy]) * S + xx;
    int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
    for (int y = 0; y < n; ++y) {
      const int w = k[y];
      a0 += static_cast<int>(s[static_cast<long long>(y) * S]) * w;
      a1 += static_cast<int>(s[plane + static_cast<long long>(y) * S]) * w;
      a2 += static_cast<int>(s[2 * plane + static_cast<long long>(y) * S]) * w;
    }
    const int xo = flip ? S - 1 - xx : xx;
    if (OUT == OUT_U8) {
      uint8_t* d = o + static_cast<long long>(y_begin + yy) * S + xo;
      d[0] = static_cast<uint8_t>(clip8(a0));
      d[static_cast<long long>(S) * S] = static_cast<uint8_t>(clip8(a1));
      d[2LL * S * S] = static_cast<uint8_t>(clip8(a2));
    } else {
      const int y = y_begin + yy;
      const long long prow = static_cast<long long>(y / pe.P) * pe.Gp + xo / pe.P;
      uint16_t* d = reinterpret_cast<uint16_t*>(o) + prow * (3LL * pe.P * pe.P) + (y % pe.P) * pe.P + xo % pe.P;
      const int acc[3] = {a0, a1, a2};
#pragma unroll
      for (int c = 0; c < 3; ++c)
        d[c * pe.P * pe.P] = to_h<OUT == OUT_PATCH_F16>(fmaf(static_cast<float>(clip8(acc[c])), pe.a[c], pe.b[c]));
    }
  }
}

template <int OUT>
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, const ViewDev* __restrict__ views, int S, int kmax,
                  uint8_t* __restrict__ out, int words_ok, const PatchEmit pe) {
  extern __shared__ __align__(16) int rs_smem[];
  const ViewDev v = views[blockIdx.y];
  const int y_begin = blockIdx.x * RBV;
  if (y_begin >= S) return;
  const int rows = min(RBV, S - y_begin);
  int* kk = rs_smem;                 // [RBV][kmax]: all lanes of a row read the same word (broadcast)
  int* ymin = kk + RBV * kmax;       // [RBV]
  int* cnt = ymin + RBV;             // [RBV]
  for (int yy = threadIdx.x; yy < rows; yy += blockDim.x) {
    int lo, n;
    axis_coeffs(v.v, v.filter, y_begin + yy + v.v.off, kk + yy * kmax, 1, lo, n);
    ymin[yy] = lo - v.row0;          // rows of the intermediate are stored from row0 on (Pillow: ybox_first)
    cnt[yy] = n;
  }
  __syncthreads();
  const long long plane = static_cast<long long>(v.n_rows) * S;
  const uint8_t* t = tmp + v.tmp_off;
  // a view = 3 S^2 bytes of planar uint8, or its Gp^2 rows of 3 P^2 16-bit patch columns (the same 3 S^2 elements)
  uint8_t* o = out + static_cast<long long>(blockIdx.y) * 3 * S * S * (OUT == OUT_U8 ? 1 : 2);
  const int ks = words_ok ? v.v.ksize : 0;
  switch (ks) {
    case 3: v_rows_fast<3, OUT>(t, plane, kk, kmax, ymin, cnt, S, y_begin, rows, v.flip, o, pe); break;
    case 5: v_rows_fast<5, OUT>(t, plane, kk, kmax, ymin, cnt, S, y_begin, rows, v.flip, o, pe); break;
    case 7: v_rows_fast<7, OUT>(t, plane, kk, kmax, ymin, cnt, S, y_begin, rows, v.flip, o, pe); break;
    case 9: v_rows_fast<9, OUT>(t, plane, kk, kmax, ymin, cnt, S, y_begin, rows, v.flip, o, pe); break;
    default: v_rows_generic<OUT>(t, plane, kk, kmax, ymin, cnt, S, y_begin, rows, v.flip, o, pe);
  }
}

// host twin of the bounds part of precompute_coeffs (same IEEE double arithmetic as the device code)
void axis_setup(int in_size, int out_size, int off, int filter, AxisDev& a) {
  const double support0 = filter == 0 ? 1.0 : 2.0;
  double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  a.scale = scale;
  a.ss = 1.0 / filterscale;
  a.support = support0 * filterscale;
  a.ksize = static_cast<int>(std::ceil(a.support)) * 2 + 1;
  a.in_size = in_size;
  a.off = off;
}
void axis_bounds(const AxisDev& a, int X, int& lo, int& hi) {
  const double center = (X + 0.5) * a.scale;
  lo = static_cast<int>(center - a.support + 0.5);
  if (lo < 0) lo = 0;
  hi = static_cast<int>(center + a.support + 0.5);
  if (hi > a.in_size) hi = a.in_size;
}

}  // namespace

size_t tta_plan(const TtaImage* images, int n_images, const TtaJob* jobs, int64_t n_jobs, int S,
                std::vector<uint8_t>* plan_bytes, int* kmax_h, int* kmax_v, int* max_rows, const char** err) {
  std::vector<ViewDev> views(static_cast<size_t>(n_jobs));
  size_t tmp_bytes = 0;
  int kh = 1, kv = 1, mr = 0;
  for (int64_t i = 0; i < n_jobs; ++i) {
    const TtaJob& j = jobs[i];
    if (j.image < 0 || j.image >= n_images) { *err = "job.image out of range"; return SIZE_MAX; }
    const TtaImage& im = images[j.image];
    if (j.crop_h < 1 || j.crop_w < 1 || j.top < 0 || j.left < 0 || j.top + j.crop_h > im.height ||
        j.left + j.crop_w > im.width) { *err = "crop box outside the source image"; return SIZE_MAX; }
    if (j.out_h < S || j.out_w < S || j.off_y < 0 || j.off_x < 0 || j.off_y + S > j.out_h || j.off_x + S > j.out_w) {
      *err = "output window outside the resized image"; return SIZE_MAX;
    }
    if (j.filter < 0 || j.filter > 1) { *err = "unknown filter"; return SIZE_MAX; }
    ViewDev& v = views[static_cast<size_t>(i)];
    v.src_off = im.offset;
    v.src_w = im.width;
    v.top = j.top;
    v.left = j.left;
    v.filter = j.filter;
    v.flip = j.flip ? 1 : 0;
    axis_setup(j.crop_w, j.out_w, j.off_x, j.filter, v.h);
    axis_setup(j.crop_h, j.out_h, j.off_y, j.filter, v.v);
    int lo0, hi0, lo1, hi1;
    axis_bounds(v.v, j.off_y, lo0, hi0);               // Pillow: ybox_first = bounds of the first output row ...
    axis_bounds(v.v, j.off_y + S - 1, lo1, hi1);       // ... ybox_last = end of the last one
    v.row0 = lo0;
    v.n_rows = hi1 - lo0;
    v.tmp_off = static_cast<long long>(tmp_bytes);
    tmp_bytes += (static_cast<size_t>(v.n_rows) * S * 3 + 255) / 256 * 256;
    kh = v.h.ksize > kh ? v.h.ksize : kh;
    kv = v.v.ksize > kv ? v.v.ksize : kv;
    mr = v.n_rows > mr ? v.n_rows : mr;
  }
  plan_bytes->resize(views.size() * sizeof(ViewDev));
  if (!views.empty()) memcpy(plan_bytes->data(), views.data(), plan_bytes->size());
  *kmax_h = kh;
  *kmax_v = kv;
  *max_rows = mr;
  return tmp_bytes;
}

cudaError_t launch_tta(const uint8_t* src, const void* views_dev, int64_t n_jobs, int S, int kmax_h, int kmax_v,
                       int max_rows, uint8_t* tmp, void* out, cudaStream_t stream, int out_mode, int patch,
                       int apply_norm) {
  if (n_jobs == 0) return cudaSuccess;
  const size_t smem_h = static_cast<size_t>(S) * (kmax_h + 2) * sizeof(int);
  const size_t smem_v = static_cast<size_t>(RBV) * (kmax_v + 2) * sizeof(int);
  if (smem_h > 200 * 1024 || smem_v > 200 * 1024 || n_jobs > 65535) return cudaErrorInvalidValue;
  if (out_mode < OUT_U8 || out_mode > OUT_PATCH_F16) return cudaErrorInvalidValue;
  PatchEmit pe{};
  if (out_mode != OUT_U8) {
    if (patch < 4 || patch % 4 != 0 || S % patch != 0) return cudaErrorInvalidValue;
    pe.P = patch;
    pe.Gp = S / patch;
    // the same expressions as rowwise.cu im2col_kernel evaluates for uint8 input: bit-identical patches
    const float mean[3] = {0.48145466f, 0.4578275f, 0.40821073f}, std[3] = {0.26862954f, 0.26130258f, 0.27577711f};
    for (int c = 0; c < 3; ++c) {
      pe.a[c] = apply_norm ? 1.0f / (255.0f * std[c]) : 1.0f / 255.0f;
      pe.b[c] = apply_norm ? -mean[c] / std[c] : 0.0f;
    }
  }
  {
    cudaError_t e = ensure_dynamic_smem(resample_h_kernel, smem_h);
    if (e == cudaSuccess) e = ensure_dynamic_smem(resample_v_kernel<OUT_U8>, smem_v);
    if (e == cudaSuccess) e = ensure_dynamic_smem(resample_v_kernel<OUT_PATCH_BF16>, smem_v);
    if (e == cudaSuccess) e = ensure_dynamic_smem(resample_v_kernel<OUT_PATCH_F16>, smem_v);
    if (e != cudaSuccess) return e;
  }
  const ViewDev* views = static_cast<const ViewDev*>(views_dev);
  dim3 gh(static_cast<unsigned>((max_rows + RBH - 1) / RBH), static_cast<unsigned>(n_jobs));
  // word-wide horizontal pass: whole aligned 32-bit words of the source are read (include/jclip_b200.h states the
  // contract: src_dev 4-byte aligned and readable up to the next multiple of 4 bytes)
  const int src_words_ok = (reinterpret_cast<uintptr_t>(src) & 3) == 0;
  resample_h_kernel<<<gh, 256, smem_h, stream>>>(src, views, S, kmax_h, tmp, src_words_ok);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 gv(static_cast<unsigned>((S + RBV - 1) / RBV), static_cast<unsigned>(n_jobs));
  // word-wide vertical pass: rows of the intermediate start on 4-byte boundaries (tmp_off is 256-byte aligned)
  const int words_ok = S % 4 == 0 && (reinterpret_cast<uintptr_t>(tmp) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0;
  uint8_t* o = static_cast<uint8_t*>(out);
  if (out_mode == OUT_U8) resample_v_kernel<OUT_U8><<<gv, 256, smem_v, stream>>>(tmp, views, S, kmax_v, o, words_ok, pe);
  else if (out_mode == OUT_PATCH_BF16) resample_v_kernel<OUT_PATCH_BF16><<<gv, 256, smem_v, stream>>>(tmp, views, S, kmax_v, o, words_ok, pe);
  else resample_v_kernel<OUT_PATCH_F16><<<gv, 256, smem_v, stream>>>(tmp, views, S, kmax_v, o, words_ok, pe);
  return cudaGetLastError();
}

}  // namespace jcb
