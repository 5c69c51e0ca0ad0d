// Generic CUDA-core evaluation of the five contraction networks (forward, input VJP, weight VJP,
// KFC, KFAC-reduce) for every geometry: arbitrary N, stride, padding, dilation, groups, and every
// dtype (fp64 included).  One tiled FMA GEMM whose operands are *gathered on the fly* with the
// index law i = s*o + d*k - p; nothing is unfolded in memory and no index pattern exists.
//
// This is the `simplify=False` / EINCONV_KERNELS_GENERIC path and the EINCONV_MATH_FP32 path (true
// fp32 / fp64 accumulation, meeting the reference's rtol 1e-5 tolerances).  The tensor-core kernels
// in tc_*.cu and the stencil kernels in depthwise.cu are the specialised fast paths.
//
// GEMM view of each network (g = group, handled by blockIdx.z; Kt = prod(k), Ot = prod(out)):
//   forward     C[(b,o), co]      = sum_{(ci,k)} X~[(b,o),(ci,k)] * W[co,(ci,k)]     expressions/convNd_forward.py:48-55
//   input VJP   C[(b,i), ci]      = sum_{(co,k)} V~[(b,i),(co,k)] * W[co,ci,k]       expressions/convNd_input_vjp.py:53-60
//   weight VJP  C[co, (ci,k)]     = sum_{(b,o)}  V[b,co,o]        * X~[(b,o),(ci,k)] expressions/convNd_weight_vjp.py:51-58
//   KFC         C[(ci,k),(ci',k')] = sum_{(b,o)} X~[(b,o),(ci,k)] * X~[(b,o),(ci',k')] expressions/convNd_kfc.py:75-90
//   KFAC-red.   u[b,(ci,k)] = sum_o X~[(b,o),(ci,k)] ;  C = sum_b u u^T              expressions/convNd_kfac_reduce.py:77-92
// where X~ is the (virtual) unfolded input and V~ the (virtual) scattered cotangent.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace einconv {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;
constexpr int PAD = 4;

// ---------------------------------------------------------------------------------------------
// Gather contexts.  ND = compile-time bound on the number of spatial dims handled in registers.
// ---------------------------------------------------------------------------------------------
template <int ND>
struct PosCtx {      // an output position (b, o..) expressed in input coordinates
  long long base;    // b * batch_stride + sum_d (s_d*o_d - p_d) * in_stride_d
  int v[ND];         // s_d * o_d - p_d
  bool valid;        // index inside the problem
};
template <int ND>
struct TapCtx {      // a (channel, tap) pair
  long long off;     // ch * channel_stride + sum_d dil_d*k_d * in_stride_d
  int dk[ND];        // dil_d * k_d
  bool valid;
};

// contiguous element strides of the input's spatial dims
template <int ND>
struct Strides {
  int s[ND];
};

template <int ND>
__device__ __forceinline__ PosCtx<ND> make_pos(const DevGeom& g, const Strides<ND>& st,
                                               long long batch_stride, int idx, int total) {
  // idx = b * out_total + oflat
  PosCtx<ND> c;
  c.valid = idx < total;
  uint32_t rem = c.valid ? (uint32_t)idx : 0u;
  long long base = 0;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      uint32_t q, o;
      g.out_div[d].divmod(rem, q, o);
      rem = q;
      c.v[d] = g.stride[d] * (int)o - g.pad[d];
      base += (long long)c.v[d] * st.s[d];
    } else {
      c.v[d] = 0;
    }
  }
  c.base = base + (long long)rem * batch_stride;  // rem == b
  return c;
}

template <int ND>
__device__ __forceinline__ TapCtx<ND> make_tap(const DevGeom& g, const Strides<ND>& st,
                                               long long chan_stride, int idx, int total) {
  // idx = ch * k_total + kflat
  TapCtx<ND> c;
  c.valid = idx < total;
  uint32_t rem = c.valid ? (uint32_t)idx : 0u;
  long long off = 0;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      uint32_t q, k;
      g.k_div[d].divmod(rem, q, k);
      rem = q;
      c.dk[d] = g.dil[d] * (int)k;
      off += (long long)c.dk[d] * st.s[d];
    } else {
      c.dk[d] = 0;
    }
  }
  c.off = off + (long long)rem * chan_stride;  // rem == channel
  return c;
}

// x~[(b,o),(ch,k)] : forward gather
template <typename T, int ND>
__device__ __forceinline__ typename Cvt<T>::acc_t gather_fwd(const T* __restrict__ x,
                                                             const DevGeom& g,
                                                             const PosCtx<ND>& p,
                                                             const TapCtx<ND>& t) {
  bool ok = p.valid && t.valid;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    if (d < g.n_dims) {
      const int i = p.v[d] + t.dk[d];
      ok = ok && (i >= 0) && (i < g.in[d]);
    }
  }
  return ok ? Cvt<T>::load(x + p.base + t.off) : typename Cvt<T>::acc_t(0);
}

// Input-VJP gather: V~[(b,i),(co,k)] = v[b, co, o] with o_d = (i_d + p_d - dil_d*k_d)/s_d if integral
template <int ND>
struct InPosCtx {   // an input position (b, i..)
  int b;
  int ip[ND];       // i_d + p_d
  bool valid;
};
template <int ND>
__device__ __forceinline__ InPosCtx<ND> make_inpos(const DevGeom& g, int idx, int total) {
  InPosCtx<ND> c;
  c.valid = idx < total;
  uint32_t rem = c.valid ? (uint32_t)idx : 0u;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      uint32_t q, i;
      g.in_div[d].divmod(rem, q, i);
      rem = q;
      c.ip[d] = (int)i + g.pad[d];
    } else {
      c.ip[d] = 0;
    }
  }
  c.b = (int)rem;
  return c;
}
template <int ND>
struct ChanTapCtx {  // (co, k..) for the input VJP
  int ch;
  int dk[ND];
  bool valid;
};
template <int ND>
__device__ __forceinline__ ChanTapCtx<ND> make_chantap(const DevGeom& g, int idx, int total) {
  ChanTapCtx<ND> c;
  c.valid = idx < total;
  uint32_t rem = c.valid ? (uint32_t)idx : 0u;
#pragma unroll
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      uint32_t q, k;
      g.k_div[d].divmod(rem, q, k);
      rem = q;
      c.dk[d] = g.dil[d] * (int)k;
    } else {
      c.dk[d] = 0;
    }
  }
  c.ch = (int)rem;
  return c;
}
template <typename T, int ND>
__device__ __forceinline__ typename Cvt<T>::acc_t gather_bwd(const T* __restrict__ v,
                                                             const DevGeom& g, int chan_base,
                                                             int chan_total,
                                                             const InPosCtx<ND>& p,
                                                             const ChanTapCtx<ND>& t) {
  bool ok = p.valid && t.valid;
  long long oflat = 0;
#pragma unroll
  for (int d = 0; d < ND; ++d) {
    if (d < g.n_dims) {
      const int num = p.ip[d] - t.dk[d];
      const int s = g.stride[d];
      const int o = (s == 1) ? num : num / s;
      ok = ok && (num >= 0) && (o * s == num) && (o < g.out[d]);
      oflat = oflat * g.out[d] + o;
    }
  }
  if (!ok) return typename Cvt<T>::acc_t(0);
  const long long off = ((long long)p.b * chan_total + chan_base + t.ch) * g.out_total + oflat;
  return Cvt<T>::load(v + off);
}

// ---------------------------------------------------------------------------------------------
// The GEMM core.  Prob supplies:
//   static constexpr bool A_RED_FAST / B_RED_FAST   thread mapping of the tile loads: true = lanes
//                                                    walk the reduction index, false = the row/col
//   int M, N, K;  int k_splits;                       problem extents; blockIdx.z = group*k_splits + split
//   RowA make_row_a(z, m) / RedA make_red_a(z, k) / acc load_a(row, red)     (same for b)
//   void store(z, m, n, acc)
// ---------------------------------------------------------------------------------------------
template <typename Prob>
__global__ void __launch_bounds__(NT) gather_gemm_kernel(const __grid_constant__ Prob prob) {
  using acc_t = typename Prob::acc_t;
  __shared__ acc_t As[2][BK][BM + PAD];
  __shared__ acc_t Bs[2][BK][BN + PAD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int z = blockIdx.z;
  const int split = z % prob.k_splits;
  const int grp = z / prob.k_splits;

  // reduction range of this split, in multiples of BK
  const int k_chunks = (prob.K + BK - 1) / BK;
  const int per_split = (k_chunks + prob.k_splits - 1) / prob.k_splits;
  const int kc_begin = split * per_split;
  const int kc_end = min(k_chunks, kc_begin + per_split);

  // ---- thread -> tile element mapping for the loads (4 A and 4 B elements per thread)
  // row-fast: row = tid % 64, red = tid / 64 + 4*j ;  red-fast: red = tid % 16, row = tid / 16 + 16*j
  int a_row[4], a_red[4], b_row[4], b_red[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if constexpr (Prob::A_RED_FAST) { a_red[j] = tid % BK; a_row[j] = tid / BK + (NT / BK) * j; }
    else                  { a_row[j] = tid % BM; a_red[j] = tid / BM + (NT / BM) * j; }
    if constexpr (Prob::B_RED_FAST) { b_red[j] = tid % BK; b_row[j] = tid / BK + (NT / BK) * j; }
    else                  { b_row[j] = tid % BN; b_red[j] = tid / BN + (NT / BN) * j; }
  }

  // contexts that stay fixed over the reduction loop
  typename Prob::RowA rowa[Prob::A_RED_FAST ? 4 : 1];
  typename Prob::RowB rowb[Prob::B_RED_FAST ? 4 : 1];
#pragma unroll
  for (int j = 0; j < (Prob::A_RED_FAST ? 4 : 1); ++j) rowa[j] = prob.make_row_a(grp, m0 + a_row[j]);
#pragma unroll
  for (int j = 0; j < (Prob::B_RED_FAST ? 4 : 1); ++j) rowb[j] = prob.make_row_b(grp, n0 + b_row[j]);

  acc_t ra[4], rb[4];
  auto fetch = [&](int kc) {
    const int k0 = kc * BK;
    if constexpr (Prob::A_RED_FAST) {
      const auto red = prob.make_red_a(grp, k0 + a_red[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = prob.load_a(rowa[j], red);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = prob.load_a(rowa[0], prob.make_red_a(grp, k0 + a_red[j]));
    }
    if constexpr (Prob::B_RED_FAST) {
      const auto red = prob.make_red_b(grp, k0 + b_red[0]);
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = prob.load_b(rowb[j], red);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = prob.load_b(rowb[0], prob.make_red_b(grp, k0 + b_red[j]));
    }
  };
  auto stash = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      As[buf][a_red[j]][a_row[j]] = ra[j];
      Bs[buf][b_red[j]][b_row[j]] = rb[j];
    }
  };

  const int tx = tid % 16;  // n direction
  const int ty = tid / 16;  // m direction
  acc_t acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = acc_t(0);

  if (kc_begin < kc_end) {
    fetch(kc_begin);
    stash(0);
    __syncthreads();
    for (int kc = kc_begin; kc < kc_end; ++kc) {
      const int buf = (kc - kc_begin) & 1;
      if (kc + 1 < kc_end) fetch(kc + 1);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        acc_t a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[buf][kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[buf][kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
      }
      if (kc + 1 < kc_end) {
        stash(buf ^ 1);
        __syncthreads();
      }
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < prob.M && n < prob.N) prob.store(grp, split, m, n, acc[i][j]);
    }
}

// ---------------------------------------------------------------------------------------------
// Problem definitions
// ---------------------------------------------------------------------------------------------
template <typename T, int ND>
struct ProbBase {
  using acc_t = typename Cvt<T>::acc_t;
  DevGeom g;
  Strides<ND> st;   // contiguous spatial strides of x
  int M, N, K, k_splits;
  int c_in_total, c_out_total;  // G*Cin_g, G*Cout_g
};

template <typename T, int ND>
static void fill_base(ProbBase<T, ND>& p, const DevGeom& g) {
  p.g = g;
  int s = 1;
  for (int d = ND - 1; d >= 0; --d) {
    if (d < g.n_dims) {
      p.st.s[d] = s;
      s *= g.in[d];
    } else {
      p.st.s[d] = 0;
    }
  }
  p.c_in_total = g.groups * g.c_in_g;
  p.c_out_total = g.groups * g.c_out_g;
  p.k_splits = 1;
}

// ---- forward --------------------------------------------------------------------------------
template <typename T, int ND>
struct FwdProb : ProbBase<T, ND> {
  using acc_t = typename Cvt<T>::acc_t;
  static constexpr bool A_RED_FAST = false;  // lanes walk output positions (contiguous in x)
  static constexpr bool B_RED_FAST = true;   // lanes walk (ci,k) (contiguous in w)
  using RowA = PosCtx<ND>;
  using RedA = TapCtx<ND>;
  struct RowB { long long off; bool valid; };
  struct RedB { int k; bool valid; };
  const T* x; const T* w; const T* bias; T* y;

  __device__ RowA make_row_a(int grp, int m) const {
    return make_pos<ND>(this->g, this->st, (long long)this->c_in_total * this->g.in_total, m, this->M);
  }
  __device__ RedA make_red_a(int grp, int k) const {
    TapCtx<ND> t = make_tap<ND>(this->g, this->st, this->g.in_total, k, this->K);
    t.off += (long long)grp * this->g.c_in_g * this->g.in_total;
    return t;
  }
  __device__ acc_t load_a(const RowA& r, const RedA& t) const { return gather_fwd<T, ND>(x, this->g, r, t); }
  __device__ RowB make_row_b(int grp, int n) const {
    return {((long long)grp * this->g.c_out_g + n) * this->K, n < this->N};
  }
  __device__ RedB make_red_b(int grp, int k) const { return {k, k < this->K}; }
  __device__ acc_t load_b(const RowB& r, const RedB& t) const {
    return (r.valid && t.valid) ? Cvt<T>::load(w + r.off + t.k) : acc_t(0);
  }
  __device__ void store(int grp, int split, int m, int n, acc_t v) const {
    const int b = m / this->g.out_total, o = m - b * this->g.out_total;
    const int co = grp * this->g.c_out_g + n;
    if (bias) v += Cvt<T>::load(bias + co);
    Cvt<T>::store(y + ((long long)b * this->c_out_total + co) * this->g.out_total + o, v);
  }
};

// ---- input VJP ------------------------------------------------------------------------------
template <typename T, int ND>
struct InVjpProb : ProbBase<T, ND> {
  using acc_t = typename Cvt<T>::acc_t;
  static constexpr bool A_RED_FAST = false;  // lanes walk input positions
  static constexpr bool B_RED_FAST = true;   // lanes walk (co,k)
  using RowA = InPosCtx<ND>;
  using RedA = ChanTapCtx<ND>;
  struct RowB { int ci; bool valid; };
  struct RedB { long long off; bool valid; };
  const T* w; const T* v; T* gx;

  __device__ RowA make_row_a(int grp, int m) const { return make_inpos<ND>(this->g, m, this->M); }
  __device__ RedA make_red_a(int grp, int k) const { return make_chantap<ND>(this->g, k, this->K); }
  __device__ acc_t load_a(const RowA& r, const RedA& t) const {
    // channel base of this group is folded in through a per-launch pointer offset (see launch)
    return gather_bwd<T, ND>(v, this->g, 0, this->c_out_total, r, t);
  }
  __device__ RowB make_row_b(int grp, int n) const { return {n, n < this->N}; }
  __device__ RedB make_red_b(int grp, int k) const {
    // k = co * Kt + kflat  ->  w[(g*Cout_g + co), ci, kflat]
    const int co = k / this->g.k_total, kf = k - co * this->g.k_total;
    return {(((long long)grp * this->g.c_out_g + co) * this->g.c_in_g) * this->g.k_total + kf, k < this->K};
  }
  __device__ acc_t load_b(const RowB& r, const RedB& t) const {
    return (r.valid && t.valid) ? Cvt<T>::load(w + t.off + (long long)r.ci * this->g.k_total) : acc_t(0);
  }
  __device__ void store(int grp, int split, int m, int n, acc_t val) const {
    const int b = m / this->g.in_total, i = m - b * this->g.in_total;
    Cvt<T>::store(gx + ((long long)b * this->c_in_total + grp * this->g.c_in_g + n) * this->g.in_total + i, val);
  }
};

// ---- weight VJP (split-K) -------------------------------------------------------------------
template <typename T, int ND>
struct WVjpProb : ProbBase<T, ND> {
  using acc_t = typename Cvt<T>::acc_t;
  static constexpr bool A_RED_FAST = true;  // lanes walk (b,o) (contiguous in v)
  static constexpr bool B_RED_FAST = true;  // lanes walk (b,o) (nearly contiguous in x)
  struct RowA { long long off; bool valid; };
  struct RedA { long long off; bool valid; };
  using RowB = TapCtx<ND>;
  using RedB = PosCtx<ND>;
  const T* x; const T* v; acc_t* partial;  // [G][splits][M][N]

  __device__ RowA make_row_a(int grp, int m) const {
    return {((long long)grp * this->g.c_out_g + m) * this->g.out_total, m < this->M};
  }
  __device__ RedA make_red_a(int grp, int k) const {
    const int b = k / this->g.out_total, o = k - b * this->g.out_total;
    return {(long long)b * this->c_out_total * this->g.out_total + o, k < this->K};
  }
  __device__ acc_t load_a(const RowA& r, const RedA& t) const {
    return (r.valid && t.valid) ? Cvt<T>::load(v + r.off + t.off) : acc_t(0);
  }
  __device__ RowB make_row_b(int grp, int n) const {
    TapCtx<ND> t = make_tap<ND>(this->g, this->st, this->g.in_total, n, this->N);
    t.off += (long long)grp * this->g.c_in_g * this->g.in_total;
    return t;
  }
  __device__ RedB make_red_b(int grp, int k) const {
    return make_pos<ND>(this->g, this->st, (long long)this->c_in_total * this->g.in_total, k, this->K);
  }
  __device__ acc_t load_b(const RowB& t, const RedB& r) const { return gather_fwd<T, ND>(x, this->g, r, t); }
  __device__ void store(int grp, int split, int m, int n, acc_t val) const {
    partial[(((long long)grp * this->k_splits + split) * this->M + m) * this->N + n] = val;
  }
};

// ---- KFC (split-K) --------------------------------------------------------------------------
template <typename T, int ND>
struct KfcProb : ProbBase<T, ND> {
  using acc_t = typename Cvt<T>::acc_t;
  static constexpr bool A_RED_FAST = true;
  static constexpr bool B_RED_FAST = true;
  using RowA = TapCtx<ND>;
  using RedA = PosCtx<ND>;
  using RowB = TapCtx<ND>;
  using RedB = PosCtx<ND>;
  const T* x; acc_t* partial;  // [G][splits][A][A]

  __device__ RowA make_row_a(int grp, int m) const {
    TapCtx<ND> t = make_tap<ND>(this->g, this->st, this->g.in_total, m, this->M);
    t.off += (long long)grp * this->g.c_in_g * this->g.in_total;
    return t;
  }
  __device__ RedA make_red_a(int grp, int k) const {
    return make_pos<ND>(this->g, this->st, (long long)this->c_in_total * this->g.in_total, k, this->K);
  }
  __device__ acc_t load_a(const RowA& t, const RedA& r) const { return gather_fwd<T, ND>(x, this->g, r, t); }
  __device__ RowB make_row_b(int grp, int n) const { return make_row_a(grp, n); }
  __device__ RedB make_red_b(int grp, int k) const { return make_red_a(grp, k); }
  __device__ acc_t load_b(const RowB& t, const RedB& r) const { return gather_fwd<T, ND>(x, this->g, r, t); }
  __device__ void store(int grp, int split, int m, int n, acc_t val) const {
    partial[(((long long)grp * this->k_splits + split) * this->M + m) * this->N + n] = val;
  }
};

// ---- u u^T for KFAC-reduce: C[g][a][a'] = sum_b u[b][g*A + a] * u[b][g*A + a'] -----------------
template <typename T, int ND>
struct OuterProb : ProbBase<T, ND> {
  using acc_t = typename Cvt<T>::acc_t;
  static constexpr bool A_RED_FAST = false;  // lanes walk a (contiguous in u)
  static constexpr bool B_RED_FAST = false;
  struct RowA { int a; bool valid; };
  struct RedA { long long off; bool valid; };
  using RowB = RowA;
  using RedB = RedA;
  const acc_t* u;   // [B][G*A]
  acc_t* partial;   // [G][1][A][A]

  __device__ RowA make_row_a(int grp, int m) const { return {grp * this->M + m, m < this->M}; }
  __device__ RedA make_red_a(int grp, int k) const {
    return {(long long)k * this->g.groups * this->M, k < this->K};
  }
  __device__ acc_t load_a(const RowA& r, const RedA& t) const {
    return (r.valid && t.valid) ? u[t.off + r.a] : acc_t(0);
  }
  __device__ RowB make_row_b(int grp, int n) const { return make_row_a(grp, n); }
  __device__ RedB make_red_b(int grp, int k) const { return make_red_a(grp, k); }
  __device__ acc_t load_b(const RowB& r, const RedB& t) const { return load_a(r, t); }
  __device__ void store(int grp, int split, int m, int n, acc_t val) const {
    partial[(((long long)grp * this->k_splits + split) * this->M + m) * this->N + n] = val;
  }
};

// ---------------------------------------------------------------------------------------------
// split-K reduction + scale + cast:  out[i] = scale * sum_s partial[grp][s][i]
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(
    const typename Cvt<T>::acc_t* __restrict__ partial, T* __restrict__ out, long long per_group,
    int k_splits, int groups, double scale) {
  using acc_t = typename Cvt<T>::acc_t;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= per_group * groups) return;
  const long long grp = idx / per_group, i = idx - grp * per_group;
  const acc_t* src = partial + grp * k_splits * per_group + i;
  acc_t s = acc_t(0);
  for (int k = 0; k < k_splits; ++k) s += src[(long long)k * per_group];
  Cvt<T>::store(out + idx, (acc_t)(s * (acc_t)scale));
}

// ---------------------------------------------------------------------------------------------
// KFAC-reduce stage 1: u[b, c, k..] = sum_{o..} x[b, c, s*o + d*k - p]   (one pass over x, HBM-bound)
//   sum_o prod_d [i_d = s_d*o_d + d_d*k_d - p_d]  ==  prod_d m_d(k_d, i_d),
//   m_d(k, i) = 1 iff t = i + p_d - d_d*k >= 0, s_d | t, t/s_d < O_d,
// so the window sums are separable: a warp reduces one input row (last dim) into K_w partial sums,
// then adds them to every outer tap the row belongs to.  One CTA per (b, c) plane.
// ---------------------------------------------------------------------------------------------
constexpr int KW_MAX = 8;

__device__ __forceinline__ bool tap_member(const DevGeom& g, int d, int i, int k) {
  if (g.member_mode) {  // transpose convolution (conv_transposeNd_kfac_reduce.py): i is an output position
    const int t = g.stride[d] * i + g.dil[d] * k - g.pad[d];
    return t >= 0 && t < g.out[d];
  }
  const int t = i + g.pad[d] - g.dil[d] * k;
  if (t < 0) return false;
  const int s = g.stride[d];
  const int o = (s == 1) ? t : t / s;
  return o * s == t && o < g.out[d];
}

template <typename T>
__global__ void __launch_bounds__(256) window_sum_kernel(const T* __restrict__ x,
                                                         typename Cvt<T>::acc_t* __restrict__ u,
                                                         const __grid_constant__ DevGeom g,
                                                         int planes, int transposed) {
  using acc_t = typename Cvt<T>::acc_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  acc_t* acc = reinterpret_cast<acc_t*>(smem_raw);  // [k_total]
  const int plane = blockIdx.x;
  if (plane >= planes) return;
  for (int t = threadIdx.x; t < g.k_total; t += blockDim.x) acc[t] = acc_t(0);
  __syncthreads();

  const int W = g.n_dims - 1;
  const int in_w = g.in[W], k_w = g.k[W];
  const int n_rows = g.in_total / in_w;
  const int k_outer = g.k_total / k_w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const T* src = x + (long long)plane * g.in_total;

  if (k_w <= KW_MAX) {
    for (int row = warp; row < n_rows; row += n_warps) {
      acc_t part[KW_MAX];
#pragma unroll
      for (int kw = 0; kw < KW_MAX; ++kw) part[kw] = acc_t(0);
      for (int iw = lane; iw < in_w; iw += 32) {
        const acc_t val = Cvt<T>::load(src + (long long)row * in_w + iw);
#pragma unroll
        for (int kw = 0; kw < KW_MAX; ++kw)
          if (kw < k_w && tap_member(g, W, iw, kw)) part[kw] += val;
      }
#pragma unroll
      for (int kw = 0; kw < KW_MAX; ++kw) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part[kw] += __shfl_xor_sync(0xffffffffu, part[kw], off);
      }
      // outer coordinates of this row, then every outer tap combination it belongs to;
      // lane l takes combinations l, l+32, ...
      int i[kMaxDims];
      {
        uint32_t rem = (uint32_t)row;
        for (int d = W - 1; d >= 0; --d) {
          uint32_t q, r;
          g.in_div[d].divmod(rem, q, r);
          i[d] = (int)r;
          rem = q;
        }
      }
      for (int kq = lane; kq < k_outer; kq += 32) {
        uint32_t krem = (uint32_t)kq;
        bool ok = true;
        for (int d = W - 1; d >= 0; --d) {
          uint32_t q, kd;
          g.k_div[d].divmod(krem, q, kd);
          krem = q;
          ok = ok && tap_member(g, d, i[d], (int)kd);
        }
        if (ok) {
#pragma unroll
          for (int kw = 0; kw < KW_MAX; ++kw)
            if (kw < k_w) atomicAdd(&acc[kq * k_w + kw], part[kw]);
        }
      }
    }
  } else {
    // wide kernels: per-pixel enumeration of taps
    for (int idx = threadIdx.x; idx < g.in_total; idx += blockDim.x) {
      uint32_t rem = (uint32_t)idx;
      int i[kMaxDims];
      for (int d = W; d >= 0; --d) {
        uint32_t q, r;
        g.in_div[d].divmod(rem, q, r);
        i[d] = (int)r;
        rem = q;
      }
      const acc_t val = Cvt<T>::load(src + idx);
      for (int kf = 0; kf < g.k_total; ++kf) {
        uint32_t krem = (uint32_t)kf;
        bool ok = true;
        for (int d = W; d >= 0; --d) {
          uint32_t q, kd;
          g.k_div[d].divmod(krem, q, kd);
          krem = q;
          ok = ok && tap_member(g, d, i[d], (int)kd);
        }
        if (ok) atomicAdd(&acc[kf], val);
      }
    }
  }
  __syncthreads();
  const int chans = g.groups * g.c_in_g;
  const int b = plane / chans, c = plane - b * chans;
  for (int t = threadIdx.x; t < g.k_total; t += blockDim.x)
    u[transposed ? ((long long)c * g.k_total + t) * g.batch + b : (long long)plane * g.k_total + t] = acc[t];
}

// ---------------------------------------------------------------------------------------------
// KFAC-reduce stage 1, fast path: one WARP per (b, c) plane, no barrier, no atomics.
// Lane l owns columns l, l+32, ... of the innermost dimension.  For every element it adds the value
// to the accumulators of the outer tap combinations its row belongs to (a warp-uniform bit mask
// looked up per row); column membership is applied once at the end when the per-lane column sums
// are reduced across the warp.  One coalesced pass over x: the kernel is HBM-bound.
//   member_d[i] = bit mask over k_d with  i = s_d*o + dil_d*k_d - p_d  for some valid o
// ---------------------------------------------------------------------------------------------
template <typename T, int MAXJ, int MAXKO>
__global__ void __launch_bounds__(256) window_sum_warp_kernel(const T* __restrict__ x,
                                                              typename Cvt<T>::acc_t* __restrict__ u,
                                                              const __grid_constant__ DevGeom g,
                                                              int planes, int transposed) {
  using acc_t = typename Cvt<T>::acc_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* member = reinterpret_cast<unsigned*>(smem_raw);  // per dim: in[d] masks, concatenated
  const int W = g.n_dims - 1;
  // membership tables (shared by the CTA's warps)
  int table_off[kMaxDims];
  {
    int off = 0;
    for (int d = 0; d <= W; ++d) {
      table_off[d] = off;
      off += g.in[d];
    }
    for (int idx = threadIdx.x; idx < off; idx += blockDim.x) {
      int d = 0;
      while (d < W && idx >= table_off[d] + g.in[d]) ++d;
      const int i = idx - table_off[d];
      unsigned mask = 0;
      for (int k = 0; k < g.k[d]; ++k)
        if (tap_member(g, d, i, k)) mask |= 1u << k;
      member[idx] = mask;
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int plane = blockIdx.x * (blockDim.x >> 5) + warp;
  if (plane >= planes) return;
  const int in_w = g.in[W], k_w = g.k[W];
  const int n_rows = g.in_total / in_w;
  const int k_outer = g.k_total / k_w;
  const T* src = x + (long long)plane * g.in_total;

  acc_t acc[MAXJ][MAXKO];
#pragma unroll
  for (int j = 0; j < MAXJ; ++j)
#pragma unroll
    for (int ko = 0; ko < MAXKO; ++ko) acc[j][ko] = acc_t(0);

  // walk the rows; outer coordinates advance like an odometer
  int coord[kMaxDims];
  for (int d = 0; d < W; ++d) coord[d] = 0;
  // R rows per iteration so that R * MAXJ independent loads are in flight per lane (one row at a
  // time left the warp latency-bound at ~0.9 TB/s)
  constexpr int R = 4;
  int kprev[kMaxDims];
  {
    int t = 1;
    for (int d = 0; d < W; ++d) {
      kprev[d] = t;
      t *= g.k[d];
    }
  }
  for (int row0 = 0; row0 < n_rows; row0 += R) {
    unsigned masks[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      // outer tap combinations this row belongs to: bit (k_0, .., k_{W-1}) flattened, last fastest
      unsigned mask = 0u;
      if (row0 + r < n_rows) {
        mask = 1u;
        for (int d = 0; d < W; ++d) {
          const unsigned md = member[table_off[d] + coord[d]];
          // mask_new[(ko_prev * K_d) + k] = mask[ko_prev] & md[k]
          unsigned out = 0;
          for (int kp = 0; kp < kprev[d]; ++kp)
            if (mask >> kp & 1u) out |= md << (kp * g.k[d]);
          mask = out;
        }
        for (int d = W - 1; d >= 0; --d) {
          if (++coord[d] < g.in[d]) break;
          coord[d] = 0;
        }
      }
      masks[r] = mask;
    }
    acc_t v[R][MAXJ];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const T* rp = src + (long long)(row0 + r) * in_w;
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = lane + 32 * j;
        v[r][j] = acc_t(0);
        if (masks[r] != 0u && c < in_w) v[r][j] = Cvt<T>::load(rp + c);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int j = 0; j < MAXJ; ++j)
#pragma unroll
        for (int ko = 0; ko < MAXKO; ++ko)
          if (masks[r] >> ko & 1u) acc[j][ko] += v[r][j];
  }

  // apply column membership and reduce across lanes
  for (int ko = 0; ko < k_outer; ++ko) {
    for (int kw = 0; kw < k_w; ++kw) {
      acc_t s = acc_t(0);
#pragma unroll
      for (int j = 0; j < MAXJ; ++j) {
        const int c = lane + 32 * j;
        acc_t a = acc_t(0);
#pragma unroll
        for (int q = 0; q < MAXKO; ++q)
          if (q == ko) a = acc[j][q];
        if (c < in_w && (member[table_off[W] + c] >> kw & 1u)) s += a;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) {
        const int t = ko * k_w + kw;
        const int chans = g.groups * g.c_in_g;
        const int b = plane / chans, c = plane - b * chans;
        u[transposed ? ((long long)c * g.k_total + t) * g.batch + b : (long long)plane * g.k_total + t] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Window sums, flat variant for planes that fit a shared table (in_total <= 12288) and <= 32 taps:
// a per-CTA table holds, for every input position of a plane, the bit mask of the taps whose window
// contains it; a warp then streams a whole plane as a flat array (consecutive lanes = consecutive
// elements, 4 x 32 loads in flight), accumulates one register per tap and reduces with shuffles.
// Persistent CTAs amortise the table over many planes.  No atomics: the result is deterministic.
// ---------------------------------------------------------------------------------------------
template <typename T, int KT>
__global__ void __launch_bounds__(256) window_sum_flat_kernel(const T* __restrict__ x,
                                                              typename Cvt<T>::acc_t* __restrict__ u,
                                                              const __grid_constant__ DevGeom g,
                                                              int planes, int transposed) {
  using acc_t = typename Cvt<T>::acc_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* table = reinterpret_cast<unsigned*>(smem_raw);
  const int n = g.in_total;
  for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
    int i[kMaxDims];
    uint32_t rem = (uint32_t)idx;
    for (int d = g.n_dims - 1; d >= 0; --d) {
      uint32_t q, r;
      g.in_div[d].divmod(rem, q, r);
      i[d] = (int)r;
      rem = q;
    }
    unsigned mask = 0;
    for (int t = 0; t < g.k_total; ++t) {
      uint32_t krem = (uint32_t)t;
      bool ok = true;
      for (int d = g.n_dims - 1; d >= 0; --d) {
        uint32_t q, kd;
        g.k_div[d].divmod(krem, q, kd);
        krem = q;
        ok = ok && tap_member(g, d, i[d], (int)kd);
      }
      if (ok) mask |= 1u << t;
    }
    table[idx] = mask;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const int chans = g.groups * g.c_in_g;
#ifndef EINCONV_WSUM_U
#define EINCONV_WSUM_U 8
#endif
  constexpr int U = EINCONV_WSUM_U;  // independent loads in flight per lane
  for (int plane = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); plane < planes;
       plane += gridDim.x * warps_per_cta) {
    const T* src = x + (long long)plane * n;
    acc_t acc[KT];
#pragma unroll
    for (int t = 0; t < KT; ++t) acc[t] = acc_t(0);
    for (int base = 0; base < n; base += 32 * U) {
      acc_t v[U];
      unsigned m[U];
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const int idx = base + q * 32 + lane;
        m[q] = idx < n ? table[idx] : 0u;
        v[q] = acc_t(0);
        if (m[q] != 0u) v[q] = Cvt<T>::load(src + idx);
      }
#pragma unroll
      for (int q = 0; q < U; ++q)
#pragma unroll
        for (int t = 0; t < KT; ++t)
          if (m[q] >> t & 1u) acc[t] += v[q];
    }
    const int b = plane / chans, c = plane - b * chans;
#pragma unroll
    for (int t = 0; t < KT; ++t) {
      if (t < g.k_total) {
        acc_t s = acc[t];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0)
          u[transposed ? ((long long)c * g.k_total + t) * g.batch + b : (long long)plane * g.k_total + t] = s;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Window sums of a 1 x 1 / stride 1 / unpadded kernel (the dense rule): every position is under the single
// tap, so u[b, c] is the plain sum of plane (b, c).  A warp streams a plane with 16-byte loads, 8 in flight
// per lane (no membership table, no predicates); planes of at most 64 elements, which are too short to
// keep a warp's loads in flight, are summed one per THREAD instead (consecutive loads of a thread hit the
// sectors its neighbours just fetched).  The flat kernel ran these layers at 3.8 TB/s.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float reg_value(float v) { return v; }
__device__ __forceinline__ double reg_value(double v) { return v; }
__device__ __forceinline__ float reg_value(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float reg_value(__half v) { return __half2float(v); }

template <typename T>
__global__ void __launch_bounds__(256) plane_sum_kernel(const T* __restrict__ x,
                                                        typename Cvt<T>::acc_t* __restrict__ u, int n, int planes,
                                                        int chans, int batch, int transposed, int vec_ok) {
  using acc_t = typename Cvt<T>::acc_t;
  constexpr int V = 16 / sizeof(T);  // elements per 16-byte load
  const int lane = threadIdx.x & 31;
  if (n <= 64) {  // one plane per thread
    for (int plane = blockIdx.x * blockDim.x + threadIdx.x; plane < planes; plane += gridDim.x * blockDim.x) {
      const T* src = x + (long long)plane * n;
      acc_t a0 = acc_t(0), a1 = acc_t(0), a2 = acc_t(0), a3 = acc_t(0);
      int i = 0;
      for (; i + 4 <= n; i += 4) {
        a0 += Cvt<T>::load(src + i);
        a1 += Cvt<T>::load(src + i + 1);
        a2 += Cvt<T>::load(src + i + 2);
        a3 += Cvt<T>::load(src + i + 3);
      }
      for (; i < n; ++i) a0 += Cvt<T>::load(src + i);
      const int b = plane / chans, c = plane - b * chans;
      u[transposed ? (long long)c * batch + b : (long long)plane] = (a0 + a1) + (a2 + a3);
    }
    return;
  }
  const int warps_per_cta = blockDim.x >> 5;
  for (int plane = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); plane < planes; plane += gridDim.x * warps_per_cta) {
    const T* src = x + (long long)plane * n;
    acc_t acc[4] = {acc_t(0), acc_t(0), acc_t(0), acc_t(0)};
    if (vec_ok) {
      const uint4* s4 = reinterpret_cast<const uint4*>(src);
      const int nv = n / V;
      constexpr int U = 8;
      for (int base = 0; base < nv; base += 32 * U) {
        uint4 v[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
          const int idx = base + q * 32 + lane;
          v[q] = idx < nv ? __ldg(s4 + idx) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
          alignas(16) T e[V];  // out-of-range chunks were loaded as zeros
          *reinterpret_cast<uint4*>(e) = v[q];
#pragma unroll
          for (int k = 0; k < V; ++k) acc[k & 3] += (acc_t)reg_value(e[k]);
        }
      }
    } else {
      for (int idx = lane; idx < n; idx += 32) acc[0] += Cvt<T>::load(src + idx);
    }
    acc_t sum = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) {
      const int b = plane / chans, c = plane - b * chans;
      u[transposed ? (long long)c * batch + b : (long long)plane] = sum;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Window sums of a 2-d plane, separable form.  Tap (kh, kw) sums x over rows(kh) x cols(kw) where both
// membership relations depend on one coordinate only, so
//     u[kh][kw] = sum_{j in cols(kw)}  ( sum_{i in rows(kh)} x[i][j] ).
// A lane owns columns j (and, for planes narrower than a warp, one of 32 / Wl row phases): per element it
// adds into at most K_h lane-local accumulators (row membership is a table bit), 3 FADD for a 3 x 3 kernel
// instead of the 9 masked adds + table look-up of the flat kernel.  The column relation is applied once
// per plane, when the K_h x K_w sums are reduced across the warp.  Loads are coalesced row segments, 4
// row steps in flight.  No atomics: deterministic.
// ---------------------------------------------------------------------------------------------
template <typename T, int KH, int Q>
__global__ void __launch_bounds__(256) window_sum_sep_kernel(const T* __restrict__ x,
                                                             typename Cvt<T>::acc_t* __restrict__ u,
                                                             const __grid_constant__ DevGeom g, int planes,
                                                             int transposed, int Wl, int rpw) {
  using acc_t = typename Cvt<T>::acc_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned* rowmask = reinterpret_cast<unsigned*>(smem_raw);  // [H]: bit kh set iff row i is under tap kh
  const int H = g.in[0], W = g.in[1], k_h = g.k[0], k_w = g.k[1];
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    unsigned m = 0;
    for (int k = 0; k < k_h; ++k)
      if (tap_member(g, 0, i, k)) m |= 1u << k;
    rowmask[i] = m;
  }
  const int lane = threadIdx.x & 31;
  const int ri = lane / Wl;        // row phase of this lane (0 when the plane is at least a warp wide)
  const int jl = lane - ri * Wl;
  unsigned cm[Q];                  // column masks of this lane's columns
  bool col_ok[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int j = jl + Wl * q;
    col_ok[q] = j < W && ri < rpw;
    unsigned m = 0;
    if (col_ok[q])
      for (int k = 0; k < k_w; ++k)
        if (tap_member(g, 1, j, k)) m |= 1u << k;
    cm[q] = m;
    col_ok[q] = col_ok[q] && m != 0u;  // columns under no tap are never read
  }
  __syncthreads();

  const int warps_per_cta = blockDim.x >> 5;
  const int chans = g.groups * g.c_in_g;
  const long long n = (long long)H * W;
  constexpr int RU = 4;  // row steps in flight
  for (int plane = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); plane < planes;
       plane += gridDim.x * warps_per_cta) {
    const T* src = x + (long long)plane * n;
    acc_t acc[Q][KH];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int k = 0; k < KH; ++k) acc[q][k] = acc_t(0);
    for (int i0 = 0; i0 < H; i0 += rpw * RU) {
      acc_t v[RU][Q];
      unsigned rm[RU];
#pragma unroll
      for (int r = 0; r < RU; ++r) {
        const int i = i0 + r * rpw + ri;
        rm[r] = i < H ? rowmask[i] : 0u;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          v[r][q] = acc_t(0);
          if (rm[r] != 0u && col_ok[q]) v[r][q] = Cvt<T>::load(src + (long long)i * W + jl + Wl * q);
        }
      }
#pragma unroll
      for (int r = 0; r < RU; ++r)
#pragma unroll
        for (int k = 0; k < KH; ++k)
          if (rm[r] >> k & 1u) {
#pragma unroll
            for (int q = 0; q < Q; ++q) acc[q][k] += v[r][q];
          }
    }
    const int b = plane / chans, c = plane - b * chans;
    acc_t mine = acc_t(0);  // lane t ends up holding tap t
#pragma unroll
    for (int kh = 0; kh < KH; ++kh) {
      if (kh >= k_h) break;
      for (int kw = 0; kw < k_w; ++kw) {
        acc_t sum = acc_t(0);
#pragma unroll
        for (int q = 0; q < Q; ++q)
          if (cm[q] >> kw & 1u) sum += acc[q][kh];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const int t = kh * k_w + kw;
        if ((t & 31) == lane) mine = sum;
        if ((t & 31) == 31 || t == g.k_total - 1) {  // flush up to 32 taps with one store instruction
          const int tt = (t & ~31) + lane;
          if (tt <= t)
            u[transposed ? ((long long)c * g.k_total + tt) * g.batch + b : (long long)plane * g.k_total + tt] = mine;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
static int pick_splits(long long tiles, int k_chunks) {
  // enough CTAs for ~3 waves, but keep >= 8 chunks per split
  long long want = (3ll * num_sms() + tiles - 1) / tiles;
  long long cap = std::max(1, k_chunks / 8);
  return (int)std::max<long long>(1, std::min<long long>(std::min(want, cap), 512));
}

template <typename Prob>
static int launch_gemm(const Prob& p, int groups, cudaStream_t stream, const char* name) {
  dim3 grid(ceil_div(p.M, BM), ceil_div(p.N, BN), groups * p.k_splits);
  if (grid.x == 0 || grid.y == 0 || grid.z == 0) return EINCONV_OK;
  EINCONV_CHECK_ARG(grid.y < 65536 && grid.z < 65536, "%s: problem too large for the generic kernel", name);
  gather_gemm_kernel<Prob><<<grid, NT, 0, stream>>>(p);
  EINCONV_CHECK_LAUNCH(name);
  return EINCONV_OK;
}

template <typename F>
static int dispatch_nd(int n_dims, F&& f) {
  if (n_dims <= 3) return f(std::integral_constant<int, 3>{});
  return f(std::integral_constant<int, kMaxDims>{});
}

static bool fits_int(long long v) { return v >= 0 && v < (1ll << 31); }

int ffma_conv_fwd(const void* x, const void* w, const void* bias, void* y, int dtype,
                  const DevGeom& g, cudaStream_t stream) {
  const long long M = (long long)g.batch * g.out_total, K = (long long)g.c_in_g * g.k_total;
  EINCONV_CHECK_ARG(fits_int(M) && fits_int(K), "conv_fwd: index space exceeds 2^31");
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    return dispatch_nd(g.n_dims, [&](auto nd) -> int {
      constexpr int ND = decltype(nd)::value;
      FwdProb<T, ND> p;
      fill_base<T, ND>(p, g);
      p.M = (int)M; p.N = g.c_out_g; p.K = (int)K;
      p.x = (const T*)x; p.w = (const T*)w; p.bias = (const T*)bias; p.y = (T*)y;
      return launch_gemm(p, g.groups, stream, "ffma_conv_fwd");
    });
  });
}

int ffma_conv_input_vjp(const void* w, const void* v, void* gx, int dtype, const DevGeom& g,
                        cudaStream_t stream) {
  const long long M = (long long)g.batch * g.in_total, K = (long long)g.c_out_g * g.k_total;
  EINCONV_CHECK_ARG(fits_int(M) && fits_int(K), "conv_input_vjp: index space exceeds 2^31");
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    return dispatch_nd(g.n_dims, [&](auto nd) -> int {
      constexpr int ND = decltype(nd)::value;
      // one launch per group: the cotangent's channel base moves with the group
      for (int grp = 0; grp < g.groups; ++grp) {
        InVjpProb<T, ND> p;
        fill_base<T, ND>(p, g);
        p.M = (int)M; p.N = g.c_in_g; p.K = (int)K;
        p.w = (const T*)w + (long long)grp * g.c_out_g * g.c_in_g * g.k_total;
        p.v = (const T*)v + (long long)grp * g.c_out_g * g.out_total;
        p.gx = (T*)gx + (long long)grp * g.c_in_g * g.in_total;
        int st = launch_gemm(p, 1, stream, "ffma_conv_input_vjp");
        if (st) return st;
      }
      return EINCONV_OK;
    });
  });
}

int64_t ffma_splitk_workspace(int op, const DevGeom& g, int dtype, int* splits_out) {
  long long M, N, K;
  if (op == EINCONV_OP_CONV_WEIGHT_VJP) {
    M = g.c_out_g; N = (long long)g.c_in_g * g.k_total;
  } else {
    M = N = (long long)g.c_in_g * g.k_total;
  }
  K = (long long)g.batch * g.out_total;
  const long long tiles = (long long)ceil_div(M, BM) * ceil_div(N, BN) * g.groups;
  const int splits = pick_splits(std::max(1ll, tiles), ceil_div(std::max(1ll, K), BK));
  if (splits_out) *splits_out = splits;
  const int64_t acc = (dtype == EINCONV_F64) ? 8 : 4;
  int64_t bytes = (int64_t)g.groups * splits * M * N * acc;
  if (op == EINCONV_OP_KFAC_REDUCE) {
    // u [B][G*A] + one partial
    bytes = ((int64_t)g.batch * g.groups * M + (int64_t)g.groups * M * N) * acc;
  }
  return bytes;
}

template <typename T>
static int launch_reduce(const void* partial, void* out, long long per_group, int splits,
                         int groups, double scale, cudaStream_t stream) {
  const long long total = per_group * groups;
  if (total == 0) return EINCONV_OK;
  splitk_reduce_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      (const typename Cvt<T>::acc_t*)partial, (T*)out, per_group, splits, groups, scale);
  EINCONV_CHECK_LAUNCH("splitk_reduce_kernel");
  return EINCONV_OK;
}

int ffma_conv_weight_vjp(const void* x, const void* v, void* gw, int dtype, const DevGeom& g,
                         void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  const long long N = (long long)g.c_in_g * g.k_total, K = (long long)g.batch * g.out_total;
  EINCONV_CHECK_ARG(fits_int(N) && fits_int(K), "conv_weight_vjp: index space exceeds 2^31");
  int splits = 1;
  const int64_t need = ffma_splitk_workspace(EINCONV_OP_CONV_WEIGHT_VJP, g, dtype, &splits);
  if (workspace_bytes < need || (need > 0 && !workspace)) {
    set_error("conv_weight_vjp: workspace of %lld bytes required, got %lld", (long long)need,
              (long long)workspace_bytes);
    return EINCONV_ERR_WORKSPACE;
  }
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    using acc_t = typename Cvt<T>::acc_t;
    return dispatch_nd(g.n_dims, [&](auto nd) -> int {
      constexpr int ND = decltype(nd)::value;
      WVjpProb<T, ND> p;
      fill_base<T, ND>(p, g);
      p.M = g.c_out_g; p.N = (int)N; p.K = (int)K; p.k_splits = splits;
      p.x = (const T*)x; p.v = (const T*)v; p.partial = (acc_t*)workspace;
      if (K == 0) {
        cudaError_t e = cudaMemsetAsync(workspace, 0, need, stream);
        if (e != cudaSuccess) return cuda_fail(e, "memset");
      } else {
        int st = launch_gemm(p, g.groups, stream, "ffma_conv_weight_vjp");
        if (st) return st;
      }
      return launch_reduce<T>(workspace, gw, (long long)p.M * p.N, splits, g.groups, 1.0, stream);
    });
  });
}

int ffma_kfc(const void* x, void* out, int dtype, const DevGeom& g, double scale, void* workspace,
             int64_t workspace_bytes, cudaStream_t stream) {
  const long long A = (long long)g.c_in_g * g.k_total, K = (long long)g.batch * g.out_total;
  EINCONV_CHECK_ARG(fits_int(A) && fits_int(K), "kfc: index space exceeds 2^31");
  int splits = 1;
  const int64_t need = ffma_splitk_workspace(EINCONV_OP_KFC, g, dtype, &splits);
  if (workspace_bytes < need || (need > 0 && !workspace)) {
    set_error("kfc: workspace of %lld bytes required, got %lld", (long long)need,
              (long long)workspace_bytes);
    return EINCONV_ERR_WORKSPACE;
  }
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    using acc_t = typename Cvt<T>::acc_t;
    return dispatch_nd(g.n_dims, [&](auto nd) -> int {
      constexpr int ND = decltype(nd)::value;
      KfcProb<T, ND> p;
      fill_base<T, ND>(p, g);
      p.M = p.N = (int)A; p.K = (int)K; p.k_splits = splits;
      p.x = (const T*)x; p.partial = (acc_t*)workspace;
      if (K == 0) {
        cudaError_t e = cudaMemsetAsync(workspace, 0, need, stream);
        if (e != cudaSuccess) return cuda_fail(e, "memset");
      } else {
        int st = launch_gemm(p, g.groups, stream, "ffma_kfc");
        if (st) return st;
      }
      return launch_reduce<T>(workspace, out, A * A, splits, g.groups, scale, stream);
    });
  });
}

// stage 1 only: u (acc_t) = window sums.  transposed = 0: [B][G*A]; 1: [G*A][B] (batch contiguous,
// the channels-first layout the tensor-core KFC kernel consumes as a 1-d "image" of length B)
int kfac_window_sums(const void* x, void* u_out, int dtype, const DevGeom& g, int transposed,
                     cudaStream_t stream) {
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    using acc_t = typename Cvt<T>::acc_t;
    acc_t* u = (acc_t*)u_out;
    const int planes = g.batch * g.groups * g.c_in_g;
    if (planes == 0 || g.k_total == 0) return EINCONV_OK;
    const int W = g.n_dims - 1;
    const int k_outer = g.k_total / g.k[W];
    long long table = 0;
    bool masks_fit = true;
    for (int d = 0; d <= W; ++d) {
      table += g.in[d];
      masks_fit = masks_fit && g.k[d] <= 32;
    }
    // 2-d planes up to 224 columns and 7 rows of taps: separable kernel (column-owning lanes)
    // dense rule: one tap that covers every position -> plain plane sums
    {
      bool dense = g.k_total == 1 && g.member_mode == 0;
      for (int d = 0; d <= W && dense; ++d)
        dense = g.stride[d] == 1 && g.pad[d] == 0 && g.out[d] == g.in[d];
      if (dense && !getenv("EINCONV_WINDOW_SUM_FLAT")) {
        const int n = g.in_total;
        const int vec_ok = ((long long)n * sizeof(T)) % 16 == 0 && ((uintptr_t)x & 15) == 0;
        const long long work_ctas = n <= 64 ? (planes + 255) / 256 : (planes + 7) / 8;
        const unsigned blocks = (unsigned)std::min<long long>(work_ctas, 16ll * num_sms());
        plane_sum_kernel<T><<<blocks, 256, 0, stream>>>((const T*)x, u, n, planes, g.groups * g.c_in_g, g.batch,
                                                       transposed, vec_ok);
        EINCONV_CHECK_LAUNCH("plane_sum_kernel");
        return EINCONV_OK;
      }
    }
    // (other 1-tap kernels keep the flat kernel)
    if (g.n_dims == 2 && g.k_total > 1 && g.in[1] <= 224 && g.k[0] <= 7 && g.k[1] <= 32 && g.in[0] <= 8192 &&
        !getenv("EINCONV_WINDOW_SUM_FLAT")) {
      const int W_ = g.in[1];
      int Wl = 32;
      while (Wl / 2 >= W_ && Wl > 1) Wl /= 2;  // smallest power of two >= W (<= 32)
      const int rpw = 32 / Wl;                  // rows covered by one warp step
      const int Q = (W_ + 31) / 32;
      const size_t smem = (size_t)g.in[0] * 4;
      // enough warps to fill the machine, persistent over planes beyond that
      const unsigned blocks = (unsigned)std::min<long long>((planes + 7) / 8, 8ll * num_sms());
      auto launch_sep = [&](auto kh_tag, auto q_tag) -> int {
        constexpr int KH = decltype(kh_tag)::value;
        constexpr int QQ = decltype(q_tag)::value;
        window_sum_sep_kernel<T, KH, QQ><<<blocks, 256, smem, stream>>>((const T*)x, u, g, planes, transposed, Wl, rpw);
        EINCONV_CHECK_LAUNCH("window_sum_sep_kernel");
        return EINCONV_OK;
      };
      auto with_q = [&](auto kh_tag) -> int {
        if (Q <= 1) return launch_sep(kh_tag, std::integral_constant<int, 1>{});
        if (Q <= 2) return launch_sep(kh_tag, std::integral_constant<int, 2>{});
        if (Q <= 4) return launch_sep(kh_tag, std::integral_constant<int, 4>{});
        return launch_sep(kh_tag, std::integral_constant<int, 7>{});
      };
      if (g.k[0] == 1) return with_q(std::integral_constant<int, 1>{});
      if (g.k[0] <= 3) return with_q(std::integral_constant<int, 3>{});
      return with_q(std::integral_constant<int, 7>{});
    }
    if (g.in_total <= 12288 && g.k_total <= 32 && !getenv("EINCONV_WINDOW_SUM_ROWS")) {
      const size_t smem = (size_t)g.in_total * 4;
      const unsigned blocks = (unsigned)std::min<long long>((planes + 7) / 8, 4ll * num_sms());
      auto launch_flat = [&](auto kt) -> int {
        constexpr int KT = decltype(kt)::value;
        auto kern = window_sum_flat_kernel<T, KT>;
        if (smem > 40 * 1024) {
          cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
          if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(window_sum_flat)");
        }
        kern<<<blocks, 256, smem, stream>>>((const T*)x, u, g, planes, transposed);
        EINCONV_CHECK_LAUNCH("window_sum_flat_kernel");
        return EINCONV_OK;
      };
      if (g.k_total == 1) return launch_flat(std::integral_constant<int, 1>{});
      if (g.k_total <= 9) return launch_flat(std::integral_constant<int, 9>{});
      return launch_flat(std::integral_constant<int, 32>{});
    }
    if (g.in[W] <= 32 * 8 && k_outer <= 9 && masks_fit && table * 4 <= 40 * 1024) {
      const unsigned blocks = (unsigned)((planes + 7) / 8);
      if (g.in[W] <= 32 * 2)
        window_sum_warp_kernel<T, 2, 9><<<blocks, 256, table * 4, stream>>>((const T*)x, u, g, planes, transposed);
      else if (g.in[W] <= 32 * 4)
        window_sum_warp_kernel<T, 4, 9><<<blocks, 256, table * 4, stream>>>((const T*)x, u, g, planes, transposed);
      else
        window_sum_warp_kernel<T, 8, 9><<<blocks, 256, table * 4, stream>>>((const T*)x, u, g, planes, transposed);
      EINCONV_CHECK_LAUNCH("window_sum_warp_kernel");
    } else {
      const size_t smem = (size_t)g.k_total * sizeof(acc_t);
      EINCONV_CHECK_ARG(smem <= 48 * 1024, "kfac_reduce: kernel too large (%d taps)", g.k_total);
      window_sum_kernel<T><<<planes, 256, smem, stream>>>((const T*)x, u, g, planes, transposed);
      EINCONV_CHECK_LAUNCH("window_sum_kernel");
    }
    return EINCONV_OK;
  });
}

int ffma_kfac_reduce(const void* x, void* out, int dtype, const DevGeom& g, double scale,
                     void* workspace, int64_t workspace_bytes, cudaStream_t stream) {
  const long long A = (long long)g.c_in_g * g.k_total;
  EINCONV_CHECK_ARG(fits_int(A), "kfac_reduce: index space exceeds 2^31");
  const int64_t need = ffma_splitk_workspace(EINCONV_OP_KFAC_REDUCE, g, dtype, nullptr);
  if (workspace_bytes < need || (need > 0 && !workspace)) {
    set_error("kfac_reduce: workspace of %lld bytes required, got %lld", (long long)need,
              (long long)workspace_bytes);
    return EINCONV_ERR_WORKSPACE;
  }
  return dispatch_dtype(dtype, [&](auto* tag) -> int {
    using T = std::remove_pointer_t<decltype(tag)>;
    using acc_t = typename Cvt<T>::acc_t;
    acc_t* u = (acc_t*)workspace;
    acc_t* partial = u + (long long)g.batch * g.groups * A;
    {
      int st1 = kfac_window_sums(x, u, dtype, g, /*transposed=*/0, stream);
      if (st1) return st1;
    }
    OuterProb<T, 3> p;
    fill_base<T, 3>(p, g);
    p.M = p.N = (int)A; p.K = g.batch; p.k_splits = 1;
    p.u = u; p.partial = partial;
    int st = launch_gemm(p, g.groups, stream, "ffma_kfac_outer");
    if (st) return st;
    return launch_reduce<T>(partial, out, A * A, 1, g.groups, scale, stream);
  });
}

}  // namespace einconv
