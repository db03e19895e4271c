// Batched GCN actor forward (include/tactor.h): 13 GCNConv layers of multimodes_actor
// (train/code/truss2D_RL.py:49-127) for B environments x N nodes at once.
//
// out = A . (X . W) + b per layer.  The rows of all environments are stacked into one [B*N, 208] activation
// matrix (hidden 200 padded to 208 with zero weights), so X . W is ONE tall GEMM per layer; the adjacency
// product is block diagonal (one N x N block per environment) and is applied to the 64-row output tile
// while it is still in shared memory, fused with bias / ReLU / the five-way sum of layer 2.
//
// Arithmetic: float32 FFMA with float32 accumulation, i.e. the reference's nominal dtype (the actor has
// no recorded outputs to pin against -- SURVEY.md section 8c "parity unpinned").
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <new>
#include <string>
#include <vector>

#include "../../include/tactor.h"
#include "../../include/tfem.h"
#include "tactor_tc.cuh"
#include <stdlib.h>

namespace tactor {

constexpr int HID = TACTOR_HIDDEN;   // 200
constexpr int LD = 208;              // padded leading dimension of every activation / weight matrix
constexpr int TM = 64;               // rows per CTA tile
constexpr int KC = 16;               // K chunk
constexpr int NTHREADS = 208;        // 26 column groups x 8 row groups, 8x8 outputs per thread

// ---------------------------------------------------------------------------------------------------------
// Y[rows, 0:208] (+)= relu( Ablk . (X[rows, 0:K] . W[K, 208]) + bias )
//   adj: [B, NODES, NODES] (adj_batched) or [NODES, NODES] shared by every environment
template <int NODES>
__global__ void __launch_bounds__(NTHREADS, 2)
gcn_layer_kernel(const float* __restrict__ X, int ldx, int K, const float* __restrict__ W,
                 const float* __restrict__ bias, const float* __restrict__ adj, int adj_batched,
                 float* __restrict__ Y, int accumulate, int M) {
  constexpr int ENVS = TM / NODES;
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                       // [KC][TM]   (transposed chunk of X)
  float* Bs = As + KC * TM;               // [KC][LD]
  float* Ts = Bs + KC * LD;               // [TM][LD]   X.W tile
  float* Ad = Ts + TM * LD;               // [ENVS][NODES(j)][NODES(i)]  adjacency, transposed
  const int tid = threadIdx.x;
  const int tx = tid % 26, ty = tid / 26;
  const int row0 = blockIdx.x * TM;

  // adjacency blocks of this tile's environments (transposed so that 8 consecutive i are contiguous)
  for (int idx = tid; idx < ENVS * NODES * NODES; idx += NTHREADS) {
    const int e = idx / (NODES * NODES), r = idx % (NODES * NODES), i = r / NODES, j = r % NODES;
    const int env = row0 / NODES + e;
    float v = 0.f;
    if (env * NODES < M) v = adj_batched ? adj[(size_t)env * NODES * NODES + r] : adj[r];
    Ad[(e * NODES + j) * NODES + i] = v;
  }

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;

  const int nchunks = (K + KC - 1) / KC;
  // register staging of the next chunk: A = 1024 floats (5 per thread), B = 832 float4 (4 per thread)
  float a_reg[5];
  float4 b_reg[4];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int idx = tid + q * NTHREADS;          // 0..1039
      const int r = idx / KC, kk = idx % KC;
      float v = 0.f;
      if (idx < TM * KC && row0 + r < M && k0 + kk < K) v = X[(size_t)(row0 + r) * ldx + k0 + kk];
      a_reg[q] = v;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int idx = tid + q * NTHREADS;          // 0..831 float4
      const int kk = idx / (LD / 4), c4 = idx % (LD / 4);
      b_reg[q] = (k0 + kk < K) ? reinterpret_cast<const float4*>(W + (size_t)(k0 + kk) * LD)[c4]
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const int idx = tid + q * NTHREADS;
      if (idx < TM * KC) As[(idx % KC) * TM + idx / KC] = a_reg[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<float4*>(Bs)[tid + q * NTHREADS] = b_reg[q];
  };

  load_chunk(0);
  for (int ch = 0; ch < nchunks; ++ch) {
    __syncthreads();                     // previous chunk fully consumed
    store_chunk();
    __syncthreads();
    if (ch + 1 < nchunks) load_chunk((ch + 1) * KC);
#pragma unroll
    for (int kk = 0; kk < KC; ++kk) {
      const float4 a0 = reinterpret_cast<const float4*>(As + kk * TM + 8 * ty)[0];
      const float4 a1 = reinterpret_cast<const float4*>(As + kk * TM + 8 * ty)[1];
      const float4 b0 = reinterpret_cast<const float4*>(Bs + kk * LD + 8 * tx)[0];
      const float4 b1 = reinterpret_cast<const float4*>(Bs + kk * LD + 8 * tx)[1];
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(a[i], b[c], acc[i][c]);
    }
  }
  // X.W tile -> shared memory
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float4* dst = reinterpret_cast<float4*>(Ts + (8 * ty + i) * LD + 8 * tx);
    dst[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    dst[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
  }
  __syncthreads();
  // adjacency product on the tile: out[i][c] = sum_j A[i][j] T[j][c] inside each environment block
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  {
    const int e = (8 * ty) / NODES, ri = (8 * ty) % NODES;
#pragma unroll 4
    for (int j = 0; j < NODES; ++j) {
      const float4 a0 = reinterpret_cast<const float4*>(Ad + (e * NODES + j) * NODES + ri)[0];
      const float4 a1 = reinterpret_cast<const float4*>(Ad + (e * NODES + j) * NODES + ri)[1];
      const float4 t0 = reinterpret_cast<const float4*>(Ts + (e * NODES + j) * LD + 8 * tx)[0];
      const float4 t1 = reinterpret_cast<const float4*>(Ts + (e * NODES + j) * LD + 8 * tx)[1];
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(a[i], t[c], acc[i][c]);
    }
  }
  // bias + ReLU (+ accumulate into the five-way sum) and store
  const float4 bs0 = reinterpret_cast<const float4*>(bias + 8 * tx)[0];
  const float4 bs1 = reinterpret_cast<const float4*>(bias + 8 * tx)[1];
  const float bb[8] = {bs0.x, bs0.y, bs0.z, bs0.w, bs1.x, bs1.y, bs1.z, bs1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row0 + 8 * ty + i;
    if (row >= M) continue;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = fmaxf(acc[i][c] + bb[c], 0.f);
    float4* dst = reinterpret_cast<float4*>(Y + (size_t)row * LD + 8 * tx);
    if (accumulate) {
      const float4 o0 = dst[0], o1 = dst[1];
      v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
      v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
    }
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

template <int NODES>
constexpr int layer_smem_bytes() {
  return (KC * TM + KC * LD + TM * LD + (TM / NODES) * NODES * NODES) * 4;
}

// ---------------------------------------------------------------------------------------------------------
// Pareto branch (truss2D_RL.py:86-95): x14 = relu(A_p (x_p W14) + b14), sum over the valid Pareto rows,
// then the reference's stack-and-reshape: x14b[b, n, h] = pooled[b, (n*200 + h) / N].  One CTA per env.
template <int NODES>
__global__ void __launch_bounds__(256)
pareto_kernel(const float* __restrict__ x_p, const float* __restrict__ A_p, const int32_t* __restrict__ n_pf,
              int P, const float* __restrict__ W14, const float* __restrict__ b14, float* __restrict__ X14b,
              float* __restrict__ pooled_out, int B) {
  __shared__ float xs[50 * 4];
  __shared__ float as[50 * 50];
  __shared__ float pooled[HID];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (b >= B) return;
  for (int i = tid; i < P * 4; i += blockDim.x) xs[i] = x_p[(size_t)b * P * 4 + i];
  for (int i = tid; i < P * P; i += blockDim.x) as[i] = A_p[(size_t)b * P * P + i];
  __syncthreads();
  const int valid = n_pf ? min(max(n_pf[b], 0), P) : P;
  if (tid < HID) {
    const float w0 = W14[0 * LD + tid], w1 = W14[1 * LD + tid], w2 = W14[2 * LD + tid], w3 = W14[3 * LD + tid];
    const float bias = b14[tid];
    float sum = 0.f;
    for (int p = 0; p < valid; ++p) {
      float u = 0.f;
      for (int q = 0; q < P; ++q) {
        const float t = fmaf(xs[q * 4 + 3], w3, fmaf(xs[q * 4 + 2], w2, fmaf(xs[q * 4 + 1], w1, xs[q * 4] * w0)));
        u = fmaf(as[p * P + q], t, u);
      }
      sum += fmaxf(u + bias, 0.f);
    }
    pooled[tid] = sum;
  }
  __syncthreads();
  if (pooled_out) {                                  // fused path: the scramble happens in the GEMM's operand generator
    for (int i = tid; i < LD; i += blockDim.x) pooled_out[(size_t)b * LD + i] = (i < HID) ? pooled[i] : 0.f;
    return;
  }
  for (int i = tid; i < NODES * LD; i += blockDim.x) {
    const int n = i / LD, h = i % LD;
    X14b[((size_t)b * NODES + n) * LD + h] = (h < HID) ? pooled[(n * HID + h) / NODES] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Output layers (truss2D_RL.py:121-125): sigmoid(A_n (x W) + b), NOUT = 2 (geo) or 3 (topo).
// One lane per node, 32 / NODES environments per warp; the A_n product runs over warp shuffles.
template <int NODES, int NOUT>
__global__ void __launch_bounds__(128)
out_layer_kernel(const float* __restrict__ X, const float* __restrict__ W, const float* __restrict__ bias,
                 const float* __restrict__ A_n, float* __restrict__ out, int M) {
  __shared__ float ws[HID * NOUT];
  __shared__ float an[NODES * NODES];
  for (int i = threadIdx.x; i < HID * NOUT; i += blockDim.x) ws[i] = W[(i / NOUT) * LD + (i % NOUT)];
  for (int i = threadIdx.x; i < NODES * NODES; i += blockDim.x) an[i] = A_n[i];
  __syncthreads();
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  float t[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) t[o] = 0.f;
  if (row < M) {
    const float4* xr = reinterpret_cast<const float4*>(X + (size_t)row * LD);
    for (int k4 = 0; k4 < HID / 4; ++k4) {
      const float4 x = xr[k4];
      const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int o = 0; o < NOUT; ++o) t[o] = fmaf(xv[u], ws[(4 * k4 + u) * NOUT + o], t[o]);
    }
  }
  const int i = lane % NODES, base = lane - i;
  float r[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) r[o] = 0.f;
  for (int j = 0; j < NODES; ++j) {
    const float a = an[i * NODES + j];
#pragma unroll
    for (int o = 0; o < NOUT; ++o) r[o] = fmaf(a, __shfl_sync(0xffffffffu, t[o], base + j), r[o]);
  }
  if (row < M) {
#pragma unroll
    for (int o = 0; o < NOUT; ++o) out[(size_t)row * NOUT + o] = 1.f / (1.f + expf(-(r[o] + bias[o])));
  }
}

// ---------------------------------------------------------------------------------------------------------
// OU noise of act() (truss2D_RL.py:41-48, 341-350): x += theta*(mu-x)*1e-4 + sigma*n, n ~ N(0,1)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void ou_noise_kernel(float* __restrict__ a, size_t n, float mu, float theta, float sigma,
                                uint64_t seed, uint64_t call, uint64_t stream_id) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t h = mix64(mix64(seed ^ (call * 0xD1342543DE82EF95ull)) + stream_id * 0x632BE59BD9B4E019ull + i);
  const float u1 = ((uint32_t)(h >> 32) + 1.0f) * 2.3283064365386963e-10f;    // (0, 1]
  const float u2 = (uint32_t)h * 2.3283064365386963e-10f;
  const float nrm = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  const float x = a[i];
  a[i] = x + (theta * (mu - x) * 0.0001f + sigma * nrm);
}

}  // namespace tactor

// =========================================================================================================
namespace {
thread_local std::string g_actor_err;
}
struct tactor_handle_s {
  int device = 0, nodes = 0, max_batch = 0;
  float* d_w[TACTOR_NLAYERS] = {};     // packed [Kpad, 208]
  float* d_b[TACTOR_NLAYERS] = {};     // [208]
  float* buf[5] = {};                  // activations [max_batch*nodes, 208]
  float* d_wimg[TACTOR_NLAYERS] = {};  // tcgen05 operand image of the hidden layers (hi/lo split, core-matrix layout)
  int* d_error = nullptr;              // set by a kernel whose mbarrier wait timed out
  bool use_tc = false;
  std::atomic<int64_t> launches{0};
  uint64_t calls = 0;
};

namespace {
int afail(int code, const std::string& msg) { g_actor_err = msg; return code; }
const int kIn[TACTOR_NLAYERS] = {13, 13, 13, 4, 200, 200, 200, 200, 200, 200, 200, 200, 200};
const int kOut[TACTOR_NLAYERS] = {200, 200, 200, 200, 200, 200, 200, 200, 200, 200, 200, 2, 3};

struct Guard {
  int prev = -1;
  explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); }
  ~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <int NODES>
cudaError_t run_forward(tactor_handle_s* h, int B, const tactor_inputs* in, float* geo, float* topo, cudaStream_t st) {
  using namespace tactor;
  const int M = B * NODES;
  const int grid = (M + TM - 1) / TM;
  const int smem = layer_smem_bytes<NODES>();
  float *x11 = h->buf[0], *x12 = h->buf[1], *x13 = h->buf[2], *x14b = h->buf[3], *S = h->buf[4];
  if (h->use_tc) {
    // fused path: Pareto embedding, then ONE kernel for the whole network (tactor_tc.cuh)
    float* pooled = h->buf[0];
    pareto_kernel<NODES><<<B, 256, 0, st>>>(in->x_p, in->A_p, in->n_pf, in->P, h->d_w[3], h->d_b[3], nullptr, pooled, B);
    tc::fused::Params p{};
    p.x_n = in->x_n; p.A_n = in->A_n; p.A_s = in->A_s; p.A_ts = in->A_n_ts; p.A_cs = in->A_n_cs; p.pooled = pooled;
    for (int k = 0; k < 3; ++k) { p.w1[k] = h->d_w[k]; p.b1[k] = h->d_b[k]; }
    for (int g2 = 0; g2 < tc::fused::NGEMM; ++g2) { p.wimg[g2] = h->d_wimg[4 + g2]; p.bias[g2] = h->d_b[4 + g2]; }
    p.w_head[0] = h->d_w[11]; p.w_head[1] = h->d_w[12]; p.b_head[0] = h->d_b[11]; p.b_head[1] = h->d_b[12];
    p.geo = geo; p.topo = topo; p.M = M; p.error_flag = h->d_error;
    tc::fused::actor_fused_kernel<NODES><<<(M + tc::TCM - 1) / tc::TCM, tc::fused::FTHREADS, tc::fused::fused_smem_bytes<NODES>(), st>>>(p);
    h->launches.fetch_add(2);
    return cudaGetLastError();
  }
  auto layer = [&](const float* X, int ldx, int K, int li, const float* adj, int batched, float* Y, int accum) {
    if (h->use_tc && h->d_wimg[li] && ldx == LD) {
      tc::gcn_layer_tc_kernel<NODES><<<(M + tc::TCM - 1) / tc::TCM, tc::THREADS, tc::smem_bytes(NODES), st>>>(
          X, K, h->d_wimg[li], h->d_b[li], adj, batched, Y, accum, M, h->d_error);
    } else {
      gcn_layer_kernel<NODES><<<grid, NTHREADS, smem, st>>>(X, ldx, K, h->d_w[li], h->d_b[li], adj, batched, Y, accum, M);
    }
    h->launches.fetch_add(1);
  };
  layer(in->x_n, 13, 13, 0, in->A_n, 0, x11, 0);
  layer(in->x_n, 13, 13, 1, in->A_n, 0, x12, 0);
  layer(in->x_n, 13, 13, 2, in->A_n, 0, x13, 0);
  pareto_kernel<NODES><<<B, 256, 0, st>>>(in->x_p, in->A_p, in->n_pf, in->P, h->d_w[3], h->d_b[3], x14b, nullptr, B);
  h->launches.fetch_add(1);
  layer(x11, LD, HID, 4, in->A_n, 0, S, 0);          // x_2_1
  layer(x12, LD, HID, 5, in->A_n_ts, 1, S, 1);       // x_2_2
  layer(x12, LD, HID, 6, in->A_n_cs, 1, S, 1);       // x_2_3
  layer(x13, LD, HID, 7, in->A_s, 1, S, 1);          // x_2_4
  layer(x14b, LD, HID, 8, in->A_n, 0, S, 1);         // x_2_5
  layer(S, LD, HID, 9, in->A_n, 0, x11, 0);          // x_3_1
  layer(S, LD, HID, 10, in->A_s, 1, x12, 0);         // x_3_2
  const int ob = 128, og = (M + ob - 1) / ob;
  out_layer_kernel<NODES, 2><<<og, ob, 0, st>>>(x11, h->d_w[11], h->d_b[11], in->A_n, geo, M);
  out_layer_kernel<NODES, 3><<<og, ob, 0, st>>>(x12, h->d_w[12], h->d_b[12], in->A_n, topo, M);
  h->launches.fetch_add(2);
  return cudaGetLastError();
}
}  // namespace

extern "C" {

const char* tactor_last_error(void) { return g_actor_err.c_str(); }

int tactor_create(const tactor_weights* w, int nodes, int max_batch, int device, tactor_handle_t* out) {
  if (!w || !out) return afail(TFEM_ERR_ARG, "null argument");
  *out = nullptr;
  if (nodes != 16 && nodes != 32) return afail(TFEM_ERR_UNSUPPORTED, "nodes must be 16 or 32");
  if (max_batch <= 0) return afail(TFEM_ERR_ARG, "max_batch must be positive");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return afail(TFEM_ERR_CUDA, "no CUDA device: no CPU path");
  if (device < 0 || device >= ndev) return afail(TFEM_ERR_ARG, "bad device index");
  tactor_handle_s* h = new (std::nothrow) tactor_handle_s();
  if (!h) return afail(TFEM_ERR_ARG, "out of host memory");
  h->device = device; h->nodes = nodes; h->max_batch = max_batch;
  Guard g(device);
  cudaError_t e = cudaSuccess;
  for (int l = 0; l < TACTOR_NLAYERS && e == cudaSuccess; ++l) {
    if (!w->kernel[l] || !w->bias[l]) { tactor_destroy(h); return afail(TFEM_ERR_ARG, "missing layer weights"); }
    const int kin = kIn[l], kout = kOut[l];
    const int kpad = (kin + tactor::KC - 1) / tactor::KC * tactor::KC;
    std::vector<float> wp((size_t)kpad * tactor::LD, 0.f), bp(tactor::LD, 0.f);
    for (int i = 0; i < kin; ++i)
      for (int o = 0; o < kout; ++o) wp[(size_t)i * tactor::LD + o] = w->kernel[l][(size_t)i * kout + o];
    for (int o = 0; o < kout; ++o) bp[o] = w->bias[l][o];
    e = cudaMalloc(&h->d_w[l], wp.size() * 4);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_b[l], bp.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_w[l], wp.data(), wp.size() * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_b[l], bp.data(), bp.size() * 4, cudaMemcpyHostToDevice);
  }
  // tcgen05 operand images of the [200,200] layers: per 32-wide K chunk, [hi|lo][kb][n(208)][4 floats]
  {
    const char* impl = getenv("TACTOR_IMPL");
    h->use_tc = !(impl && strcmp(impl, "ffma") == 0);
  }
  for (int l = 4; l <= 10 && e == cudaSuccess && h->use_tc; ++l) {
    const int K = kIn[l], kout = kOut[l];
    std::vector<float> img;
    for (int c = 0; c * tactor::tc::KCH < K; ++c) {
      const int kw = tactor::tc::chunk_kw(K, c), nkb = kw / 4;
      for (int part = 0; part < 2; ++part)
        for (int kb = 0; kb < nkb; ++kb)
          for (int n = 0; n < tactor::tc::TCN; ++n)
            for (int t = 0; t < 4; ++t) {
              const int k = c * tactor::tc::KCH + 4 * kb + t;
              float v = (n < kout && k < K) ? w->kernel[l][(size_t)k * kout + n] : 0.f;
              uint32_t bits;
              memcpy(&bits, &v, 4);
              bits &= 0xFFFFE000u;
              float hi;
              memcpy(&hi, &bits, 4);
              img.push_back(part == 0 ? hi : v - hi);
            }
    }
    e = cudaMalloc(&h->d_wimg[l], img.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(h->d_wimg[l], img.data(), img.size() * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMalloc(&h->d_error, 4);
  if (e == cudaSuccess) e = cudaMemset(h->d_error, 0, 4);
  if (e == cudaSuccess && h->use_tc) {
    if (nodes == 16) e = cudaFuncSetAttribute(tactor::tc::fused::actor_fused_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tactor::tc::fused::fused_smem_bytes<16>());
    else e = cudaFuncSetAttribute(tactor::tc::fused::actor_fused_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tactor::tc::fused::fused_smem_bytes<32>());
  }
  const size_t rows = ((size_t)max_batch * nodes + tactor::tc::TCM - 1) / tactor::tc::TCM * tactor::tc::TCM;
  for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaMalloc(&h->buf[i], rows * tactor::LD * 4);
  if (e == cudaSuccess) {
    if (nodes == 16) e = cudaFuncSetAttribute(tactor::gcn_layer_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, tactor::layer_smem_bytes<16>());
    else e = cudaFuncSetAttribute(tactor::gcn_layer_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, tactor::layer_smem_bytes<32>());
  }
  if (e != cudaSuccess) { tactor_destroy(h); return afail(TFEM_ERR_CUDA, std::string("actor setup: ") + cudaGetErrorString(e)); }
  *out = h;
  return TFEM_OK;
}

int tactor_destroy(tactor_handle_t h) {
  if (!h) return TFEM_OK;
  Guard g(h->device);
  for (int l = 0; l < TACTOR_NLAYERS; ++l) { if (h->d_w[l]) cudaFree(h->d_w[l]); if (h->d_b[l]) cudaFree(h->d_b[l]); }
  for (int i = 0; i < 5; ++i) if (h->buf[i]) cudaFree(h->buf[i]);
  for (int l = 0; l < TACTOR_NLAYERS; ++l) if (h->d_wimg[l]) cudaFree(h->d_wimg[l]);
  if (h->d_error) cudaFree(h->d_error);
  delete h;
  return TFEM_OK;
}

static int check_inputs(tactor_handle_t h, int B, const tactor_inputs* in, float* geo, float* topo) {
  if (!h || !in || !geo || !topo) return afail(TFEM_ERR_ARG, "null argument");
  if (B < 0 || B > h->max_batch) return afail(TFEM_ERR_ARG, "batch exceeds max_batch");
  if (!in->x_n || !in->A_n || !in->A_s || !in->A_n_ts || !in->A_n_cs || !in->x_p || !in->A_p)
    return afail(TFEM_ERR_ARG, "x_n, A_n, A_s, A_n_ts, A_n_cs, x_p and A_p are required");
  if (in->P < 1 || in->P > 50) return afail(TFEM_ERR_ARG, "P must be in 1..50 (MAX_FRONT)");
  return TFEM_OK;
}

int tactor_forward(tactor_handle_t h, int B, const tactor_inputs* in, float* geo, float* topo, void* stream) {
  if (int rc = check_inputs(h, B, in, geo, topo)) return rc;
  if (B == 0) return TFEM_OK;
  Guard g(h->device);
  cudaError_t e = (h->nodes == 16) ? run_forward<16>(h, B, in, geo, topo, (cudaStream_t)stream)
                                   : run_forward<32>(h, B, in, geo, topo, (cudaStream_t)stream);
  if (e != cudaSuccess) return afail(TFEM_ERR_CUDA, std::string("actor forward: ") + cudaGetErrorString(e));
  return TFEM_OK;
}

int tactor_act(tactor_handle_t h, int B, const tactor_inputs* in, float* geo, float* topo, float mu, float theta,
               float sigma, uint64_t seed, void* stream) {
  if (int rc = tactor_forward(h, B, in, geo, topo, stream)) return rc;
  if (B == 0 || (sigma == 0.f && theta == 0.f)) return TFEM_OK;
  Guard g(h->device);
  const size_t ng = (size_t)B * h->nodes * 2, nt = (size_t)B * h->nodes * 3;
  const uint64_t call = h->calls++;
  tactor::ou_noise_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, (cudaStream_t)stream>>>(geo, ng, mu, theta, sigma, seed, call, 1);
  tactor::ou_noise_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, (cudaStream_t)stream>>>(topo, nt, mu, theta, sigma, seed, call, 2);
  h->launches.fetch_add(2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return afail(TFEM_ERR_CUDA, std::string("ou noise: ") + cudaGetErrorString(e));
  return TFEM_OK;
}

int64_t tactor_launch_count(tactor_handle_t h) { return h ? h->launches.load() : 0; }

}  // extern "C"
