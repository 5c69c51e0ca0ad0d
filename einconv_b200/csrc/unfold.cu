// K1: unfoldNd (im2col) and its adjoint fold (col2im) for arbitrary N, stride, dilation, padding.
//
// Replaces the einsum of einconv/expressions/convNd_unfold.py:42-53 (three one-hot GEMMs in the
// reference) by a gather:  out[b, c, k.., o..] = x[b, c, s*o + d*k - p] or 0.
//
// Work decomposition (input-stationary): a CTA owns `cb` consecutive (b, c) planes, ONE virtual
// coordinate of every outer spatial dimension (all but the last two; "virtual" = shifted into the
// zero-padded frame so that padding planes are visited too) and a tile of output rows of the
// second-to-last dimension.  It stages exactly the input pixels those outputs can touch into
// shared memory once (zero-filled where the virtual coordinate is padding, compacted by
// gcd(stride, dilation) so never-read rows / columns are neither loaded nor stored), then streams
// every kernel tap that maps onto this plane to global memory as contiguous runs: for a fixed tap
// the tile's outputs are `th * O_w` consecutive elements of `out`, written by consecutive lanes.
// x is therefore read once from L2/HBM (plus row halos) while `out` (prod(k) times larger) is
// written with full-line coalesced stores — the kernel is bound by HBM write bandwidth.
#include <algorithm>
#include <type_traits>

#include "common.cuh"

namespace einconv {

struct UnfoldParams {
  DevGeom g;  // n_dims >= 2 (a dummy unit dimension is inserted in front of a 1-d problem)
  long long xs_b, xs_c;       // element strides of x: batch, channel
  long long xs[kMaxDims];     // element strides of the spatial dims
  int channels;               // C = groups * c_in_g
  int n_outer;                // n_dims - 2
  int vrange[kMaxDims];       // virtual extent of outer dim d
  int v_total;                // prod(vrange[0..n_outer))
  int k_outer_total;          // prod(k[0..n_outer))
  int k_inner_total;          // k[H] * k[W]
  int tile_h, n_tiles;        // output rows per tile, tiles per plane
  int cb;                     // (b, c) planes per CTA
  int n_bc_groups;            // ceil(B*C / cb)
  int gh, gw;                 // gcd(stride, dilation) of the two inner dims
  int sh, dh, sw, dw;         // stride/g and dilation/g of the inner dims
  int rows_cap;               // staged rows for a full tile
  int cols;                   // staged columns
  FastDiv ow_div;             // / O_w
  FastDiv kw_div;             // / K_w
  FastDiv kinner_div;         // / (K_h * K_w)
  int n_seg, seg_shift;       // segments each (plane, tap) run is split into (power of two)
};

static __host__ __device__ inline int igcd(int a, int b) {
  while (b) {
    int t = a % b;
    a = b;
    b = t;
  }
  return a;
}

template <typename T>
__global__ void __launch_bounds__(256) unfold_plane_kernel(const T* __restrict__ x,
                                                           T* __restrict__ out,
                                                           const __grid_constant__ UnfoldParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* plane = reinterpret_cast<T*>(smem_raw);

  const DevGeom& g = p.g;
  const int H = g.n_dims - 2, W = g.n_dims - 1;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int n_warps = blockDim.x >> 5;

  // ---- decode CTA coordinates: tile fastest, then virtual outer coordinate, then plane group
  uint32_t bid = blockIdx.x;
  const int tile = bid % p.n_tiles;
  bid /= p.n_tiles;
  uint32_t vflat = bid % p.v_total;
  const int bcg = bid / p.v_total;

  int v[kMaxDims];  // real input coordinate of outer dims (may be outside [0, in) = padding)
  bool plane_real = true;
  long long x_outer_off = 0;
  for (int d = p.n_outer - 1; d >= 0; --d) {
    const int vd = (int)(vflat % p.vrange[d]) - g.pad[d];
    vflat /= p.vrange[d];
    v[d] = vd;
    if (vd < 0 || vd >= g.in[d]) plane_real = false;
    else x_outer_off += (long long)vd * p.xs[d];
  }

  // ---- which outer tap combinations land on this plane?  (block-uniform, usually 0..3 of them)
  // Early exit if none: e.g. odd planes when stride and dilation are both even.
  int n_valid = 0;
  for (int kq = 0; kq < p.k_outer_total; ++kq) {
    int rem = kq;
    bool ok = true;
    for (int d = p.n_outer - 1; d >= 0; --d) {
      const int kd = rem % g.k[d];
      rem /= g.k[d];
      const int t = v[d] + g.pad[d] - g.dil[d] * kd;
      if (t < 0 || t % g.stride[d] != 0 || t / g.stride[d] >= g.out[d]) ok = false;
    }
    n_valid += ok;
  }
  if (n_valid == 0) return;

  const int oh0 = tile * p.tile_h;
  const int th = min(p.tile_h, g.out[H] - oh0);
  const int rows = (p.sh * (th - 1) + p.dh * (g.k[H] - 1)) + 1;  // staged rows (compacted by gh)
  const int cols = p.cols;
  const int plane_elems = p.rows_cap * cols;
  const int bc0 = bcg * p.cb;
  const int n_planes = min(p.cb, g.batch * p.channels - bc0);

  // ---- stage: plane[pl][jr][jc] = x[b, c, v.., r0 + jr*gh, -p_w + jc*gw] or 0
  {
    const int r0 = g.stride[H] * oh0 - g.pad[H];
    for (int ri = warp; ri < n_planes * rows; ri += n_warps) {
      const int pl = ri / rows;
      const int jr = ri - pl * rows;
      const int r = r0 + jr * p.gh;
      const bool row_real = plane_real && r >= 0 && r < g.in[H];
      const int bc = bc0 + pl;
      const int b = bc / p.channels;
      const int c = bc - b * p.channels;
      const T* src = x + (long long)b * p.xs_b + (long long)c * p.xs_c + x_outer_off +
                     (long long)r * p.xs[H];
      T* dst = plane + pl * plane_elems + jr * cols;
      for (int jc = lane; jc < cols; jc += 32) {
        const int cc = jc * p.gw - g.pad[W];
        T val = T(0);
        if (row_real && cc >= 0 && cc < g.in[W]) val = __ldg(src + (long long)cc * p.xs[W]);
        dst[jc] = val;
      }
    }
  }
  __syncthreads();

  // ---- emit: for every valid outer tap combo, every inner tap, the contiguous run of th*O_w outputs
  const int chunk = th * g.out[W];
  const int n_wchunks = (chunk + 31) >> 5;
  const long long out_plane = (long long)g.k_total * g.out_total;  // elements per (b, c)
  const int inner_out = g.out[H] * g.out[W];

  for (int kq = 0; kq < p.k_outer_total; ++kq) {
    int rem = kq;
    bool ok = true;
    int o_outer_flat = 0, o_mul = 1;
    for (int d = p.n_outer - 1; d >= 0; --d) {
      const int kd = rem % g.k[d];
      rem /= g.k[d];
      const int t = v[d] + g.pad[d] - g.dil[d] * kd;
      if (t < 0 || t % g.stride[d] != 0 || t / g.stride[d] >= g.out[d]) ok = false;
      else {
        o_outer_flat += (t / g.stride[d]) * o_mul;
        o_mul *= g.out[d];
      }
    }
    if (!ok) continue;

    // out[bc][kq * k_inner + tap][o_outer_flat * inner_out + oh0 * O_w + e]
    const long long base = (long long)kq * p.k_inner_total * g.out_total +
                           (long long)o_outer_flat * inner_out + (long long)oh0 * g.out[W];
    // work item = (plane, tap, segment of 32-wide chunks); a warp walks its segment's chunks
    const int per_seg = (n_wchunks + p.n_seg - 1) >> p.seg_shift;
    const int n_items = (n_planes * p.k_inner_total) << p.seg_shift;
    for (int wi = warp; wi < n_items; wi += n_warps) {
      const int seg = wi & (p.n_seg - 1);
      uint32_t pl, tap, kh, kw;
      p.kinner_div.divmod((uint32_t)(wi >> p.seg_shift), pl, tap);
      p.kw_div.divmod(tap, kh, kw);
      const T* src = plane + pl * plane_elems + p.dh * (int)kh * cols + p.dw * (int)kw;
      T* dst = out + (long long)(bc0 + (int)pl) * out_plane + base + (long long)tap * g.out_total;
      const int wc_end = min((seg + 1) * per_seg, n_wchunks);
#pragma unroll 4
      for (int wc = seg * per_seg; wc < wc_end; ++wc) {
        const int e = (wc << 5) + lane;
        if (e < chunk) {
          uint32_t ohl, ow;
          p.ow_div.divmod((uint32_t)e, ohl, ow);
          dst[e] = src[p.sh * (int)ohl * cols + p.sw * (int)ow];
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fold: gx[bc, i..] = sum_k gu[bc, k.., o..(i,k)]   (gather form of the adjoint)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fold_kernel(const T* __restrict__ gu, T* __restrict__ gx,
                                                   const __grid_constant__ DevGeom g,
                                                   long long total) {
  using acc_t = typename Cvt<T>::acc_t;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long bc = idx / g.in_total;
  uint32_t rem = (uint32_t)(idx - bc * g.in_total);
  int i[kMaxDims];
  for (int d = g.n_dims - 1; d >= 0; --d) {
    uint32_t q, r;
    g.in_div[d].divmod(rem, q, r);
    i[d] = (int)r;
    rem = q;
  }
  const T* src = gu + bc * (long long)g.k_total * g.out_total;
  acc_t acc = acc_t(0);
  for (int kf = 0; kf < g.k_total; ++kf) {
    uint32_t krem = (uint32_t)kf;
    bool ok = true;
    int oflat = 0, omul = 1;
    for (int d = g.n_dims - 1; d >= 0; --d) {
      uint32_t q, kd;
      g.k_div[d].divmod(krem, q, kd);
      krem = q;
      const int t = i[d] + g.pad[d] - g.dil[d] * (int)kd;
      if (t < 0 || t % g.stride[d] != 0 || t / g.stride[d] >= g.out[d]) ok = false;
      else {
        oflat += (t / g.stride[d]) * omul;
        omul *= g.out[d];
      }
    }
    if (ok) acc += Cvt<T>::load(src + (long long)kf * g.out_total + oflat);
  }
  Cvt<T>::store(gx + idx, acc);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
template <typename T>
static int launch_unfold_t(const void* x, void* out, const UnfoldParams& p, size_t smem,
                           unsigned grid, cudaStream_t stream) {
  auto kern = unfold_plane_kernel<T>;
  if (smem > 48 * 1024) {
    cudaError_t err =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return cuda_fail(err, "cudaFuncSetAttribute(unfold)");
  }
  kern<<<grid, 256, smem, stream>>>((const T*)x, (T*)out, p);
  EINCONV_CHECK_LAUNCH("unfold_plane_kernel");
  return EINCONV_OK;
}

int launch_unfold(const void* x, const int64_t* x_strides, void* out, int dtype,
                  const einconv_geom* geom, cudaStream_t stream) {
  DevGeom g0;
  int st = make_dev_geom(geom, &g0, /*need_c_out=*/false);
  if (st) return st;
  const int64_t esz = elem_size(dtype);
  EINCONV_CHECK_ARG(esz > 0, "unfold: unknown dtype %d", dtype);

  UnfoldParams p{};
  const int N0 = g0.n_dims;
  const int C = g0.groups * g0.c_in_g;
  p.channels = C;

  // contiguous default strides
  long long cs[kMaxDims + 2];
  {
    long long s = 1;
    for (int d = N0 - 1; d >= 0; --d) {
      cs[d + 2] = s;
      s *= g0.in[d];
    }
    cs[1] = s;
    cs[0] = s * C;
  }
  const long long* xs = x_strides ? (const long long*)x_strides : cs;
  p.xs_b = xs[0];
  p.xs_c = xs[1];

  // normalise to >= 2 spatial dims by prepending a unit dimension
  const int shift = (N0 == 1) ? 1 : 0;
  EINCONV_CHECK_ARG(N0 + shift <= kMaxDims, "unfold: too many dimensions (%d)", N0);
  DevGeom& g = p.g;
  g = g0;
  g.n_dims = N0 + shift;
  if (shift) {
    g.in[0] = 1; g.k[0] = 1; g.stride[0] = 1; g.dil[0] = 1; g.pad[0] = 0; g.out[0] = 1;
    p.xs[0] = 0;
    for (int d = 0; d < N0; ++d) {
      g.in[d + 1] = g0.in[d]; g.k[d + 1] = g0.k[d]; g.stride[d + 1] = g0.stride[d];
      g.dil[d + 1] = g0.dil[d]; g.pad[d + 1] = g0.pad[d]; g.out[d + 1] = g0.out[d];
      p.xs[d + 1] = xs[d + 2];
    }
  } else {
    for (int d = 0; d < N0; ++d) p.xs[d] = xs[d + 2];
  }
  const int N = g.n_dims, H = N - 2, W = N - 1;

  const long long out_elems = (long long)g.batch * C * g.k_total * (long long)g.out_total;
  if (out_elems == 0) return EINCONV_OK;  // empty output: nothing to write

  p.n_outer = N - 2;
  long long v_total = 1, k_outer = 1;
  for (int d = 0; d < p.n_outer; ++d) {
    p.vrange[d] = g.stride[d] * (g.out[d] - 1) + g.dil[d] * (g.k[d] - 1) + 1;
    v_total *= p.vrange[d];
    k_outer *= g.k[d];
  }
  EINCONV_CHECK_ARG(v_total < (1ll << 30) && k_outer < (1ll << 20), "unfold: geometry too large");
  p.v_total = (int)v_total;
  p.k_outer_total = (int)k_outer;
  p.k_inner_total = g.k[H] * g.k[W];

  p.gh = igcd(g.stride[H], g.dil[H]);
  p.gw = igcd(g.stride[W], g.dil[W]);
  p.sh = g.stride[H] / p.gh; p.dh = g.dil[H] / p.gh;
  p.sw = g.stride[W] / p.gw; p.dw = g.dil[W] / p.gw;
  p.cols = p.sw * (g.out[W] - 1) + p.dw * (g.k[W] - 1) + 1;
  p.ow_div = FastDiv((uint32_t)g.out[W]);
  p.kw_div = FastDiv((uint32_t)g.k[W]);

  // tile of output rows: stage at most ~40 KB per plane
  const long long budget = 40 * 1024;
  auto rows_for = [&](int th) { return (long long)p.sh * (th - 1) + (long long)p.dh * (g.k[H] - 1) + 1; };
  int th = g.out[H];
  while (th > 1 && rows_for(th) * p.cols * esz > budget) th = (th + 1) / 2;
  EINCONV_CHECK_ARG(rows_for(th) * p.cols * esz <= 200 * 1024,
                    "unfold: one row tile does not fit in shared memory");
  p.tile_h = th;
  p.n_tiles = ceil_div(g.out[H], th);
  p.rows_cap = (int)rows_for(th);

  // planes per CTA: aim at >= 16k output elements per CTA, <= 48 KB of staging
  const long long per_plane_out = (long long)p.k_inner_total * th * g.out[W];
  const long long per_plane_smem = (long long)p.rows_cap * p.cols * esz;
  long long cb = (16384 + per_plane_out - 1) / per_plane_out;
  cb = std::min<long long>(cb, std::max<long long>(1, (48 * 1024) / per_plane_smem));
  cb = std::max<long long>(1, std::min<long long>(cb, (long long)g.batch * C));
  cb = std::min<long long>(cb, 64);
  p.cb = (int)cb;
  p.n_bc_groups = ceil_div((long long)g.batch * C, cb);
  p.kinner_div = FastDiv((uint32_t)p.k_inner_total);
  {
    // split each (plane, tap) run so that the CTA's 8 warps get >= 4 items each
    const long long runs = cb * p.k_inner_total;
    const int wchunks = ceil_div((long long)th * g.out[W], 32);
    int shift_ = 0;
    while ((runs << shift_) < 32 && (1 << (shift_ + 1)) <= wchunks) ++shift_;
    p.seg_shift = shift_;
    p.n_seg = 1 << shift_;
  }

  const long long grid = (long long)p.n_bc_groups * p.v_total * p.n_tiles;
  EINCONV_CHECK_ARG(grid > 0 && grid < (1ll << 31), "unfold: grid too large");
  const size_t smem = (size_t)(cb * per_plane_smem);

  switch (esz) {
    case 2: return launch_unfold_t<uint16_t>(x, out, p, smem, (unsigned)grid, stream);
    case 4: return launch_unfold_t<uint32_t>(x, out, p, smem, (unsigned)grid, stream);
    default: return launch_unfold_t<unsigned long long>(x, out, p, smem, (unsigned)grid, stream);
  }
}

int launch_fold(const void* gu, void* gx, int dtype, const einconv_geom* geom,
                cudaStream_t stream) {
  DevGeom g;
  int st = make_dev_geom(geom, &g, /*need_c_out=*/false);
  if (st) return st;
  const long long total = (long long)g.batch * g.groups * g.c_in_g * g.in_total;
  if (total == 0) return EINCONV_OK;
  const long long blocks = (total + 255) / 256;
  EINCONV_CHECK_ARG(blocks < (1ll << 31), "fold: tensor too large");
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    fold_kernel<T><<<(unsigned)blocks, 256, 0, stream>>>((const T*)gu, (T*)gx, g, total);
    EINCONV_CHECK_LAUNCH("fold_kernel");
    return EINCONV_OK;
  });
}

}  // namespace einconv
