// Host-resident rollout step (include/trollout.h): pieces of the batch flow through three streams so that the
// PCIe upload, the two kernels-with-C-ABI (tactor_act, tfem_step) and the PCIe download overlap.
#include <cuda_runtime.h>
#include <stdint.h>
#include <new>
#include <string>
#include <vector>

#include "../../include/trollout.h"

namespace {
struct Dev {
  float *x_n = nullptr, *A_s = nullptr, *A_ts = nullptr, *A_cs = nullptr, *raw_n = nullptr, *raw_e = nullptr, *mr = nullptr;
  float *point = nullptr, *a_geo = nullptr, *a_topo = nullptr, *x_p = nullptr, *A_p = nullptr, *A_n = nullptr;
  int32_t *status = nullptr, *n_pf = nullptr;
  uint8_t* coin = nullptr;
};
constexpr int PMAX = 50;
}  // namespace

struct trollout_handle_s {
  tfem_handle_t env = nullptr;
  tactor_handle_t actor = nullptr;
  int device = 0, max_batch = 0, piece = 0;
  tfem_dims dims{};
  Dev d;
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> ev_in, ev_run;
  std::vector<void*> allocs;
};

namespace {
thread_local std::string g_roll_err;
int rfail(int code, const std::string& m) { g_roll_err = m; return code; }

template <typename T>
cudaError_t dalloc(trollout_handle_s* h, T** p, size_t count) {
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T) + 256);
  if (e == cudaSuccess) h->allocs.push_back(*p);
  return e;
}
}  // namespace

extern "C" {

const char* trollout_last_error(void) { return g_roll_err.c_str(); }

int trollout_create(tfem_handle_t env, tactor_handle_t actor, int max_batch, int pieces, trollout_handle_t* out) {
  if (!env || !actor || !out) return rfail(TFEM_ERR_ARG, "null argument");
  *out = nullptr;
  if (max_batch <= 0 || pieces <= 0) return rfail(TFEM_ERR_ARG, "max_batch and pieces must be positive");
  trollout_handle_s* h = new (std::nothrow) trollout_handle_s();
  if (!h) return rfail(TFEM_ERR_ARG, "out of host memory");
  h->env = env; h->actor = actor; h->max_batch = max_batch;
  if (tfem_get_dims(env, &h->dims) != TFEM_OK) { delete h; return rfail(TFEM_ERR_ARG, "bad env handle"); }
  int piece = (max_batch + pieces - 1) / pieces;
  h->piece = (piece + 31) / 32 * 32;               // piece boundaries on 32 environments: every sub-array stays 16-byte aligned
  h->device = h->dims.device;
  if (h->device < 0) { delete h; return rfail(TFEM_ERR_CUDA, "tables-only env handle: libtfem has no CPU path"); }
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  cudaError_t e = cudaSetDevice(h->device);
  const size_t B = (size_t)max_batch, N = h->dims.N, E = h->dims.E;
  if (e == cudaSuccess) e = dalloc(h, &h->d.x_n, B * N * 13);
  if (e == cudaSuccess) e = dalloc(h, &h->d.A_s, B * N * N);
  if (e == cudaSuccess) e = dalloc(h, &h->d.A_ts, B * N * N);
  if (e == cudaSuccess) e = dalloc(h, &h->d.A_cs, B * N * N);
  if (e == cudaSuccess) e = dalloc(h, &h->d.raw_n, B * N * 12);
  if (e == cudaSuccess) e = dalloc(h, &h->d.raw_e, B * E * 21);
  if (e == cudaSuccess) e = dalloc(h, &h->d.mr, B * N * 2);
  if (e == cudaSuccess) e = dalloc(h, &h->d.point, B * 4);
  if (e == cudaSuccess) e = dalloc(h, &h->d.a_geo, B * N * 2);
  if (e == cudaSuccess) e = dalloc(h, &h->d.a_topo, B * N * 3);
  if (e == cudaSuccess) e = dalloc(h, &h->d.x_p, B * PMAX * 4);
  if (e == cudaSuccess) e = dalloc(h, &h->d.A_p, B * PMAX * PMAX);
  if (e == cudaSuccess) e = dalloc(h, &h->d.A_n, N * N);
  if (e == cudaSuccess) e = dalloc(h, &h->d.status, B);
  if (e == cudaSuccess) e = dalloc(h, &h->d.n_pf, B);
  if (e == cudaSuccess) e = dalloc(h, &h->d.coin, B);
  if (e == cudaSuccess) {
    std::vector<float> an(N * N);
    if (tfem_get_table(env, TFEM_TAB_A_N, an.data(), an.size() * 4) != TFEM_OK) e = cudaErrorInvalidValue;
    else e = cudaMemcpy(h->d.A_n, an.data(), an.size() * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_in, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_run, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->s_out, cudaStreamNonBlocking);
  const int npieces = (max_batch + h->piece - 1) / h->piece;
  for (int i = 0; i < npieces && e == cudaSuccess; ++i) {
    cudaEvent_t a, b;
    e = cudaEventCreateWithFlags(&a, cudaEventDisableTiming);
    if (e == cudaSuccess) { h->ev_in.push_back(a); e = cudaEventCreateWithFlags(&b, cudaEventDisableTiming); }
    if (e == cudaSuccess) h->ev_run.push_back(b);
  }
  if (prev_dev >= 0) cudaSetDevice(prev_dev);
  if (e != cudaSuccess) {
    trollout_destroy(h);
    return rfail(TFEM_ERR_CUDA, std::string("rollout setup: ") + cudaGetErrorString(e));
  }
  *out = h;
  return TFEM_OK;
}

int trollout_destroy(trollout_handle_t h) {
  if (!h) return TFEM_OK;
  for (cudaEvent_t ev : h->ev_in) cudaEventDestroy(ev);
  for (cudaEvent_t ev : h->ev_run) cudaEventDestroy(ev);
  if (h->s_in) cudaStreamDestroy(h->s_in);
  if (h->s_run) cudaStreamDestroy(h->s_run);
  if (h->s_out) cudaStreamDestroy(h->s_out);
  for (void* p : h->allocs) cudaFree(p);
  delete h;
  return TFEM_OK;
}

int trollout_bytes_per_env(trollout_handle_t h, int P, size_t* h2d, size_t* d2h) {
  if (!h) return rfail(TFEM_ERR_ARG, "null argument");
  const size_t N = h->dims.N, E = h->dims.E;
  const size_t state = 4 * (N * 13 + 3 * N * N + N * 12 + E * 21 + N * 2);
  if (h2d) *h2d = state + 1 + 4 * ((size_t)P * 4 + (size_t)P * P);
  if (d2h) *d2h = state + 4 * 4 + 4 + 4 * (N * 2 + N * 3);
  return TFEM_OK;
}

int trollout_step_host(trollout_handle_t h, int B, const trollout_io* io, float mu, float theta, float sigma,
                       uint64_t seed) {
  if (!h || !io) return rfail(TFEM_ERR_ARG, "null argument");
  if (B < 0 || B > h->max_batch) return rfail(TFEM_ERR_ARG, "batch exceeds max_batch");
  if (B == 0) return TFEM_OK;
  const trollout_state &si = io->in, &so = io->out;
  if (!si.x_n || !si.A_s || !si.A_n_ts || !si.A_n_cs || !si.nN_x_n || !si.nN_x_e || !si.move_range || !io->x_p || !io->A_p)
    return rfail(TFEM_ERR_ARG, "the parent state tuple, x_p and A_p are required");
  if (io->P < 1 || io->P > PMAX) return rfail(TFEM_ERR_ARG, "P must be in 1..50");
  const size_t N = h->dims.N, E = h->dims.E, P = (size_t)io->P;
  const Dev& d = h->d;
  int prev_dev = -1;
  cudaGetDevice(&prev_dev);
  cudaError_t e = (prev_dev == h->device) ? cudaSuccess : cudaSetDevice(h->device);
  struct Restore { int dev; ~Restore() { if (dev >= 0) cudaSetDevice(dev); } } restore{prev_dev == h->device ? -1 : prev_dev};
  auto up = [&](void* dst, const void* src, size_t bytes) {
    if (e == cudaSuccess && src) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->s_in);
  };
  auto down = [&](void* dst, const void* src, size_t bytes) {
    if (e == cudaSuccess && dst) e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->s_out);
  };
  int piece_idx = 0;
  for (int lo = 0; lo < B && e == cudaSuccess; lo += h->piece, ++piece_idx) {
    const size_t l = (size_t)lo, nb = (size_t)((lo + h->piece <= B) ? h->piece : (B - lo));
    // ---- upload ----
    up(d.x_n + l * N * 13, si.x_n + l * N * 13, nb * N * 13 * 4);
    up(d.A_s + l * N * N, si.A_s + l * N * N, nb * N * N * 4);
    up(d.A_ts + l * N * N, si.A_n_ts + l * N * N, nb * N * N * 4);
    up(d.A_cs + l * N * N, si.A_n_cs + l * N * N, nb * N * N * 4);
    up(d.raw_n + l * N * 12, si.nN_x_n + l * N * 12, nb * N * 12 * 4);
    up(d.raw_e + l * E * 21, si.nN_x_e + l * E * 21, nb * E * 21 * 4);
    up(d.mr + l * N * 2, si.move_range + l * N * 2, nb * N * 2 * 4);
    up(d.x_p + l * P * 4, io->x_p + l * P * 4, nb * P * 4 * 4);
    up(d.A_p + l * P * P, io->A_p + l * P * P, nb * P * P * 4);
    if (io->coin) up(d.coin + l, io->coin + l, nb);
    if (io->n_pf) up(d.n_pf + l, io->n_pf + l, nb * 4);
    if (e == cudaSuccess) e = cudaEventRecord(h->ev_in[piece_idx], h->s_in);
    // ---- act + step ----
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_run, h->ev_in[piece_idx], 0);
    if (e != cudaSuccess) break;
    tactor_inputs ai{};
    ai.x_n = d.x_n + l * N * 13; ai.A_n = d.A_n; ai.A_s = d.A_s + l * N * N; ai.A_n_ts = d.A_ts + l * N * N;
    ai.A_n_cs = d.A_cs + l * N * N; ai.x_p = d.x_p + l * P * 4; ai.A_p = d.A_p + l * P * P;
    ai.n_pf = io->n_pf ? d.n_pf + l : nullptr; ai.P = io->P;
    int rc = tactor_act(h->actor, (int)nb, &ai, d.a_geo + l * N * 2, d.a_topo + l * N * 3, mu, theta, sigma, seed, h->s_run);
    if (rc != TFEM_OK) return rfail(rc, std::string("rollout: ") + tactor_last_error());
    tfem_step_in in{};
    in.set_node = d.raw_n + l * N * 12; in.set_element = d.raw_e + l * E * 21; in.a_geo = d.a_geo + l * N * 2;
    in.a_topo = d.a_topo + l * N * 3; in.coin = io->coin ? d.coin + l : nullptr; in.move_range = d.mr + l * N * 2;
    tfem_step_out o{};
    o.x_n = d.x_n + l * N * 13; o.A_s = d.A_s + l * N * N; o.A_n_ts = d.A_ts + l * N * N; o.A_n_cs = d.A_cs + l * N * N;
    o.nN_x_n = d.raw_n + l * N * 12; o.nN_x_e = d.raw_e + l * E * 21; o.point = d.point + l * 4; o.status = d.status + l;
    rc = tfem_step(h->env, (int)nb, &in, &o, h->s_run);
    if (rc != TFEM_OK) return rfail(rc, std::string("rollout: ") + tfem_last_error());
    e = cudaEventRecord(h->ev_run[piece_idx], h->s_run);
    // ---- download ----
    if (e == cudaSuccess) e = cudaStreamWaitEvent(h->s_out, h->ev_run[piece_idx], 0);
    down(so.x_n ? so.x_n + l * N * 13 : nullptr, d.x_n + l * N * 13, nb * N * 13 * 4);
    down(so.A_s ? so.A_s + l * N * N : nullptr, d.A_s + l * N * N, nb * N * N * 4);
    down(so.A_n_ts ? so.A_n_ts + l * N * N : nullptr, d.A_ts + l * N * N, nb * N * N * 4);
    down(so.A_n_cs ? so.A_n_cs + l * N * N : nullptr, d.A_cs + l * N * N, nb * N * N * 4);
    down(so.nN_x_n ? so.nN_x_n + l * N * 12 : nullptr, d.raw_n + l * N * 12, nb * N * 12 * 4);
    down(so.nN_x_e ? so.nN_x_e + l * E * 21 : nullptr, d.raw_e + l * E * 21, nb * E * 21 * 4);
    down(so.move_range ? so.move_range + l * N * 2 : nullptr, d.mr + l * N * 2, nb * N * 2 * 4);
    down(io->point ? io->point + l * 4 : nullptr, d.point + l * 4, nb * 4 * 4);
    down(io->status ? io->status + l : nullptr, d.status + l, nb * 4);
    down(io->a_geo ? io->a_geo + l * N * 2 : nullptr, d.a_geo + l * N * 2, nb * N * 2 * 4);
    down(io->a_topo ? io->a_topo + l * N * 3 : nullptr, d.a_topo + l * N * 3, nb * N * 3 * 4);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_out);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->s_in);
  if (e != cudaSuccess) return rfail(TFEM_ERR_CUDA, std::string("rollout step: ") + cudaGetErrorString(e));
  return TFEM_OK;
}

}  // extern "C"
