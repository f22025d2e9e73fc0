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
//                       coalesced, writes the planar [3, S, S] uint8 view (mirrored if asked)
#include <cmath>
#include <cstdint>
#include <cstring>
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
// Writes the taps to k[0..ksize) (zero padded), returns (xmin, n) through the references.
__device__ __forceinline__ void axis_coeffs(const AxisDev& a, int kind, int X, int* k, int& xmin, int& n) {
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
    k[x] = w < 0.0 ? __double2int_rz(__dadd_rn(-0.5, f)) : __double2int_rz(__dadd_rn(0.5, f));
  }
  for (int x = cnt; x < a.ksize; ++x) k[x] = 0;
  xmin = lo;
  n = cnt;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= PREC;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// ---------------------------------------------------------------------------------- horizontal pass
__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, const ViewDev* __restrict__ views, int S, int kmax,
                  uint8_t* __restrict__ tmp) {
  extern __shared__ __align__(16) int rs_smem[];
  const ViewDev v = views[blockIdx.y];
  const int r_begin = blockIdx.x * RBH;
  if (r_begin >= v.n_rows) return;
  int* kk = rs_smem;                 // [S][kmax]
  int* xmin = kk + S * kmax;         // [S]
  int* cnt = xmin + S;               // [S]
  for (int xx = threadIdx.x; xx < S; xx += blockDim.x) {
    int lo, n;
    axis_coeffs(v.h, v.filter, xx + v.h.off, kk + xx * kmax, lo, n);
    xmin[xx] = lo;
    cnt[xx] = n;
  }
  __syncthreads();
  const int r_end = min(r_begin + RBH, v.n_rows);
  const long long pitch = 3LL * v.src_w;
  const long long plane = static_cast<long long>(v.n_rows) * S;
  for (int p = threadIdx.x; p < (r_end - r_begin) * S; p += blockDim.x) {
    const int r = r_begin + p / S, xx = p % S;
    const uint8_t* s = src + v.src_off + (v.top + v.row0 + r) * pitch + 3LL * (v.left + xmin[xx]);
    const int* k = kk + xx * kmax;
    const int n = cnt[xx];
    int a0 = 1 << (PREC - 1), a1 = a0, a2 = a0;
    for (int x = 0; x < n; ++x) {
      const int w = k[x];
      a0 += static_cast<int>(__ldg(s + 3 * x + 0)) * w;
      a1 += static_cast<int>(__ldg(s + 3 * x + 1)) * w;
      a2 += static_cast<int>(__ldg(s + 3 * x + 2)) * w;
    }
    uint8_t* d = tmp + v.tmp_off + static_cast<long long>(r) * S + xx;
    d[0] = clip8(a0);
    d[plane] = clip8(a1);
    d[2 * plane] = clip8(a2);
  }
}

// ---------------------------------------------------------------------------------- vertical pass
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ tmp, const ViewDev* __restrict__ views, int S, int kmax,
                  uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) int rs_smem[];
  const ViewDev v = views[blockIdx.y];
  const int y_begin = blockIdx.x * RBV;
  if (y_begin >= S) return;
  const int rows = min(RBV, S - y_begin);
  int* kk = rs_smem;                 // [RBV][kmax]
  int* ymin = kk + RBV * kmax;       // [RBV]
  int* cnt = ymin + RBV;             // [RBV]
  for (int yy = threadIdx.x; yy < rows; yy += blockDim.x) {
    int lo, n;
    axis_coeffs(v.v, v.filter, y_begin + yy + v.v.off, kk + yy * kmax, lo, n);
    ymin[yy] = lo - v.row0;          // rows of the intermediate are stored from row0 on (Pillow: ybox_first)
    cnt[yy] = n;
  }
  __syncthreads();
  const long long plane = static_cast<long long>(v.n_rows) * S;
  const uint8_t* t = tmp + v.tmp_off;
  uint8_t* o = out + static_cast<long long>(blockIdx.y) * 3 * S * S;
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
    const int xo = v.flip ? S - 1 - xx : xx;           // RandomHorizontalFlip acts on the finished S x S crop
    uint8_t* d = o + static_cast<long long>(y_begin + yy) * S + xo;
    d[0] = clip8(a0);
    d[static_cast<long long>(S) * S] = clip8(a1);
    d[2LL * S * S] = clip8(a2);
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
                       int max_rows, uint8_t* tmp, uint8_t* out, cudaStream_t stream) {
  if (n_jobs == 0) return cudaSuccess;
  const size_t smem_h = static_cast<size_t>(S) * (kmax_h + 2) * sizeof(int);
  const size_t smem_v = static_cast<size_t>(RBV) * (kmax_v + 2) * sizeof(int);
  if (smem_h > 200 * 1024 || smem_v > 200 * 1024 || n_jobs > 65535) return cudaErrorInvalidValue;
  static size_t attr_h = 0, attr_v = 0;
  if (smem_h > 48 * 1024 && smem_h > attr_h) {
    cudaError_t e = cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_h));
    if (e != cudaSuccess) return e;
    attr_h = smem_h;
  }
  if (smem_v > 48 * 1024 && smem_v > attr_v) {
    cudaError_t e = cudaFuncSetAttribute(resample_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_v));
    if (e != cudaSuccess) return e;
    attr_v = smem_v;
  }
  const ViewDev* views = static_cast<const ViewDev*>(views_dev);
  dim3 gh(static_cast<unsigned>((max_rows + RBH - 1) / RBH), static_cast<unsigned>(n_jobs));
  resample_h_kernel<<<gh, 256, smem_h, stream>>>(src, views, S, kmax_h, tmp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  dim3 gv(static_cast<unsigned>((S + RBV - 1) / RBV), static_cast<unsigned>(n_jobs));
  resample_v_kernel<<<gv, 256, smem_v, stream>>>(tmp, views, S, kmax_v, out);
  return cudaGetLastError();
}

}  // namespace jcb
