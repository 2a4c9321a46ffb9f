// K3 (SIMT form) -- SAGE projections + relation combine + ReLU in strict fp32 FMA arithmetic.
//
// Replaces lin_l(mean) + lin_r(x_dst) of each SAGEConv (train_gnn.py:177-184,194-197) and the
// combine relu(1.0*direct + 0.75*social) / relu(post_update) (train_gnn.py:187-198):
//     out = act( sum_i alpha_i * A_i @ W_i^T + bias )
// treated as ONE GEMM over the concatenated K of all terms, so every activation row is read
// once and no intermediate [N,H] tensor is written.  This is the shape-generic path (any K that
// keeps rows 16-byte aligned, any N); the tcgen05 tile path in proj_tc.cu takes the large
// regular shapes.
#include "common.cuh"

namespace trg {
namespace {

constexpr int kThreads = 256;
constexpr int BM = 64, BN = 64, BK = 16;

struct ProjArgs {
  const void* a[4];
  const void* w[4];
  int k[4];
  float alpha[4];
  int n_terms;
  const float* bias;
  const float* row_scale;
  int64_t n_rows;
  int hidden;
  int relu;
  void* out;
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float* f);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float* f) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) proj_simt(const ProjArgs p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Ws[BK][BN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row lr, k offset lk
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < p.n_terms; ++t) {
    const T* A = reinterpret_cast<const T*>(p.a[t]);
    const T* W = reinterpret_cast<const T*>(p.w[t]);
    const int K = p.k[t];
    const float alpha = p.alpha[t];
    for (int k0 = 0; k0 < K; k0 += BK) {
      float fa[4] = {0.f, 0.f, 0.f, 0.f}, fw[4] = {0.f, 0.f, 0.f, 0.f};
      if (m0 + lr < p.n_rows && k0 + lk < K) load4<T>(A + (m0 + lr) * K + k0 + lk, fa);
      if (n0 + lr < p.hidden && k0 + lk < K) load4<T>(W + (int64_t)(n0 + lr) * K + k0 + lk, fw);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        As[lk + i][lr] = alpha * fa[i];
        Ws[lk + i][lr] = fw[i];
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        const float w4[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], w4[j], acc[i][j]);
      }
    }
  }
  const int c0 = n0 + tx * 4;
  if (c0 >= p.hidden) return;
  float b[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = p.bias[c0 + j];
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = m0 + ty * 4 + i;
    if (r >= p.n_rows) continue;
    float o[4];
    const float rs = p.row_scale ? p.row_scale[r] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      o[j] = acc[i][j] * rs + b[j];
      if (p.relu) o[j] = fmaxf(o[j], 0.f);
    }
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + r * p.hidden + c0) =
          make_float4(o[0], o[1], o[2], o[3]);
    } else {
      uint2 v;
      v.x = Elem<__nv_bfloat16>::pack2(o[0], o[1]);
      v.y = Elem<__nv_bfloat16>::pack2(o[2], o[3]);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + r * p.hidden + c0) = v;
    }
  }
}

}  // namespace

int proj_simt_launch(const trg_proj_term* terms, int n_terms, const float* bias,
                     const float* row_scale, int64_t n_rows, int hidden, int dtype, int relu, void* out,
                     cudaStream_t st) {
  ProjArgs p{};
  p.row_scale = row_scale;
  for (int i = 0; i < n_terms; ++i) {
    p.a[i] = terms[i].a; p.w[i] = terms[i].w; p.k[i] = terms[i].k; p.alpha[i] = terms[i].alpha;
  }
  p.n_terms = n_terms; p.bias = bias; p.n_rows = n_rows; p.hidden = hidden; p.relu = relu; p.out = out;
  dim3 grid((unsigned)ceil_div<int64_t>(n_rows, BM), (unsigned)ceil_div(hidden, BN));
  if (dtype == TRG_F32)
    proj_simt<float><<<grid, kThreads, 0, st>>>(p);
  else
    proj_simt<__nv_bfloat16><<<grid, kThreads, 0, st>>>(p);
  count_launch();
  TRG_LAUNCH_OK();
  return TRG_OK;
}

}  // namespace trg
