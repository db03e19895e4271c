/* TEST INFRASTRUCTURE ONLY -- C restatement of the reference truss env-step for whole batches.
 *
 * Same algorithm, same order of floating-point operations as oracle/truss_oracle.py (the pinned Python restatement, whose
 * header lists the reference file:line of every stage); this file exists so that the parity suite can check EVERY
 * environment of a BASELINE.json-sized batch in seconds, and so that bench.py has a compiled multi-core CPU baseline for
 * the FEM-only workloads.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it; the product (libtfem.so) never does.  Pinned: tests/test_c_oracle.py compares it with the golden transitions
 * recorded from the reference itself (bit-exact heights, weak flags, sections, move range, point; 1e-12 on the FP64
 * fields) and with the Python oracle on random walks.
 *
 *   action decode, supports, round to cm ... test/0x/code/truss2D_ENV.py:379-431
 *   three constraint passes ................. :434-455
 *   symmetry copy, min(section) ............. small :460-553 / large :460-673
 *   move range .............................. truss2D_GEN.py:118-133
 *   element stiffness, assembly, solve ...... FEM_2Dtruss.py:284-337 (dense, free-DOF order; LU with partial pivoting like
 *                                             LAPACK dgesv behind np.linalg.solve)
 *   member forces, yield ratio, energy ...... FEM_2Dtruss.py:341-431
 *   objectives / constraint point ........... truss2D_ENV.py:566-589
 *
 * Numeric model: a height is (value, weak).  weak = the reference holds a python int/float; two weak
 * operands compute in float64, anything else is rounded to float32 first and computed in float32.  Build with
 * -ffp-contract=off (no FMA contraction) and without -ffast-math. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 32
#define MAXE 76
#define MAXDOF 64
#define NSEC 5

typedef struct {
  int32_t N, E, ndof, truss_type;            /* truss_type: 0 bridge, 1 roof */
  const int32_t* conn;                       /* [E][2] */
  const int32_t* tnsc;                       /* [N][2] 1-based DOF ids, free first */
  const int32_t* res;                        /* [N][2] */
  const int32_t* top;                        /* [N] */
  const int32_t* pair;                       /* [N] */
  const int32_t* sym_src;                    /* [2][N]: [coin][i] = source node of y[i] */
  const int32_t* sym_pairs;                  /* [npairs][2] */
  int32_t npairs, pad;
  const double* x;                           /* [N] */
  const double* target;                      /* [N] tar_y on top nodes (unused elsewhere) */
  const double* P;                           /* [ndof] load vector */
  const double* sec_area;                    /* [NSEC] m^2 */
  double y_min, y_max, d_min, max_def, young, allow;
  float int_obj1, int_obj2;
} oracle_family;

typedef struct {                             /* per-environment outputs; any pointer may be NULL */
  double* y;        /* [B][N] */
  uint8_t* weak;    /* [B][N] */
  int32_t* section; /* [B][E] */
  float* move_range;/* [B][N][2] new range (NOT the in/out stale one) */
  double* d;        /* [B][ndof] */
  double* axial;    /* [B][E] */
  double* ratio;    /* [B][E] */
  int32_t* iscompress; /* [B][E] */
  double* U;        /* [B] */
  float* point;     /* [B][4] */
  int32_t* status;  /* [B] 1 = singular stiffness (the reference raises LinAlgError) */
} oracle_out;

typedef struct { double v; int weak; } TS;
static inline double f32(double x) { return (double)(float)x; }
static inline TS W(double v) { TS t = {v, 1}; return t; }
static inline TS S(double v) { TS t = {f32(v), 0}; return t; }
static inline TS ts_add(TS a, TS b) { if (a.weak && b.weak) return W(a.v + b.v); TS t = {f32(f32(a.v) + f32(b.v)), 0}; return t; }
static inline TS ts_sub(TS a, TS b) { if (a.weak && b.weak) return W(a.v - b.v); TS t = {f32(f32(a.v) - f32(b.v)), 0}; return t; }
static inline TS ts_mul(TS a, TS b) { if (a.weak && b.weak) return W(a.v * b.v); TS t = {f32(f32(a.v) * f32(b.v)), 0}; return t; }
static inline TS ts_abs(TS a) { a.v = fabs(a.v); return a; }
static inline int ts_lt(TS a, TS b) { return (a.weak && b.weak) ? (a.v < b.v) : (f32(a.v) < f32(b.v)); }
static inline int ts_gt(TS a, TS b) { return (a.weak && b.weak) ? (a.v > b.v) : (f32(a.v) > f32(b.v)); }

/* np.argmax on floats: first maximum, a NaN wins and stops the scan */
static int np_argmax(const double* v, int n) {
  double mp = v[0];
  int idx = 0;
  if (mp != mp) return 0;
  for (int i = 1; i < n; ++i)
    if (!(v[i] <= mp)) { mp = v[i]; idx = i; if (mp != mp) break; }
  return idx;
}
/* np.float32.__round__(2): rint(x*100)/100 in float32 */
static double round2_f32(double v) { return f32(rint(f32(v * 100.0)) / 100.0); }

/* np.sum of a contiguous float32 vector, n <= 128 (numpy pairwise_sum with 8 accumulators) */
static double pairwise_sum_f32(const double* a, int n) {
  if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r = f32(r + a[i]); return r; }
  double r[8];
  for (int k = 0; k < 8; ++k) r[k] = a[k];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] = f32(r[k] + a[i + k]);
  double res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])));
  for (; i < n; ++i) res = f32(res + a[i]);
  return res;
}

static void clip01(float* a, int n) {        /* truss2D_ENV.py:379-391, NaN passes through */
  for (int i = 0; i < n; ++i) { if (a[i] > 1.f) a[i] = 1.f; else if (a[i] < 0.f) a[i] = 0.f; }
}

static void move_range(const oracle_family* f, const TS* y, TS* up, TS* down) {
  for (int i = 0; i < f->N; ++i) {
    TS yp = y[f->pair[i]];
    if (f->top[i] == 1) {
      up[i] = ts_abs(ts_sub(W(f->y_max), y[i]));
      down[i] = ts_abs(ts_sub(ts_sub(y[i], yp), W(f->d_min)));
    } else if (f->truss_type == 0) {
      up[i] = W(0); down[i] = W(0);
    } else {
      up[i] = ts_abs(ts_sub(ts_sub(yp, y[i]), W(f->d_min)));
      down[i] = ts_abs(ts_sub(y[i], W(f->y_min)));
    }
  }
}

/* dense LU with partial pivoting; returns 1 if exactly singular */
static int lu_solve(double* K, double* b, int n) {
  for (int c = 0; c < n; ++c) {
    int piv = c;
    double best = fabs(K[c * n + c]);
    for (int r = c + 1; r < n; ++r) { double v = fabs(K[r * n + c]); if (v > best) { best = v; piv = r; } }
    if (!(best > 0.0)) return 1;
    if (piv != c) {
      for (int k = 0; k < n; ++k) { double t = K[c * n + k]; K[c * n + k] = K[piv * n + k]; K[piv * n + k] = t; }
      double t = b[c]; b[c] = b[piv]; b[piv] = t;
    }
    const double inv = 1.0 / K[c * n + c];
    for (int r = c + 1; r < n; ++r) {
      const double l = K[r * n + c] * inv;
      if (l != 0.0) {
        for (int k = c + 1; k < n; ++k) K[r * n + k] -= l * K[c * n + k];
        b[r] -= l * b[c];
      }
    }
  }
  for (int r = n - 1; r >= 0; --r) {
    double s = b[r];
    for (int k = r + 1; k < n; ++k) s -= K[r * n + k] * b[k];
    b[r] = s / K[r * n + r];
  }
  return 0;
}

/* FEM + objectives of one environment on typed heights y and sections sec; writes outputs of environment b */
static void evaluate(const oracle_family* f, const TS* y, const int* sec, const oracle_out* o, size_t b) {
  const int N = f->N, E = f->E, n = f->ndof;
  double K[MAXDOF * MAXDOF], Kc[MAXDOF * MAXDOF], d[MAXDOF];
  double L[MAXE], c[MAXE], s[MAXE], k[MAXE], A[MAXE];
  memset(K, 0, sizeof(double) * n * n);
  for (int e = 0; e < E; ++e) {
    const int n0 = f->conn[2 * e], n1 = f->conn[2 * e + 1];
    const double dx = f->x[n1] - f->x[n0], dy = y[n1].v - y[n0].v;
    L[e] = sqrt(dx * dx + dy * dy);
    c[e] = dx / L[e]; s[e] = dy / L[e];
    A[e] = f->sec_area[sec[e]];
    k[e] = f->young * A[e] / L[e];
    const double kg[4][4] = {
        {k[e] * c[e] * c[e], k[e] * c[e] * s[e], -k[e] * c[e] * c[e], -k[e] * c[e] * s[e]},
        {k[e] * c[e] * s[e], k[e] * s[e] * s[e], -k[e] * c[e] * s[e], -k[e] * s[e] * s[e]},
        {-k[e] * c[e] * c[e], -k[e] * c[e] * s[e], k[e] * c[e] * c[e], k[e] * c[e] * s[e]},
        {-k[e] * c[e] * s[e], -k[e] * s[e] * s[e], k[e] * c[e] * s[e], k[e] * s[e] * s[e]}};
    const int ids[4] = {f->tnsc[2 * n0], f->tnsc[2 * n0 + 1], f->tnsc[2 * n1], f->tnsc[2 * n1 + 1]};
    for (int p = 0; p < 4; ++p)
      for (int q = 0; q < 4; ++q)
        if (ids[p] <= n && ids[q] <= n) K[(ids[p] - 1) * n + ids[q] - 1] += kg[p][q];
  }
  memcpy(Kc, K, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) d[i] = f->P[i];
  const int singular = lu_solve(Kc, d, n);
  if (o->status) o->status[b] = singular;
  if (singular) for (int i = 0; i < n; ++i) d[i] = NAN;
  double node_d[MAXN][2];
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < 2; ++j) node_d[i][j] = (f->tnsc[2 * i + j] <= n) ? d[f->tnsc[2 * i + j] - 1] : 0.0;
  if (o->d) memcpy(o->d + b * n, d, sizeof(double) * n);
  if (o->U) {                                /* U = 1/2 d^T K d (gen_U_full) */
    double u = 0.0;
    for (int r = 0; r < n; ++r) { double t = 0.0; for (int q = 0; q < n; ++q) t += K[r * n + q] * d[q]; u += d[r] * t; }
    o->U[b] = 0.5 * u;
  }
  double all_s[MAXE], all_v[MAXE], all_d[MAXN], all_dt[MAXN];
  for (int e = 0; e < E; ++e) {
    const int n0 = f->conn[2 * e], n1 = f->conn[2 * e + 1];
    const double u0 = c[e] * node_d[n0][0] + s[e] * node_d[n0][1];
    const double u2 = c[e] * node_d[n1][0] + s[e] * node_d[n1][1];
    const double q0 = k[e] * u0 + (-k[e]) * u2;
    const double ratio = fabs(q0 / A[e]) / f->allow;
    if (o->axial) o->axial[b * E + e] = q0;
    if (o->ratio) o->ratio[b * E + e] = ratio;
    if (o->iscompress) o->iscompress[b * E + e] = (q0 <= 0) ? 0 : 1;
    all_s[e] = f32(ratio);
    all_v[e] = f32(A[e] * L[e]);
  }
  for (int i = 0; i < N; ++i) {
    all_d[i] = 0.0; all_dt[i] = 0.0;
    if (f->top[i] == 1) all_dt[i] = f32(ts_abs(ts_sub(W(f->target[i]), y[i])).v);
    else all_d[i] = f32(node_d[i][1] / f->max_def);
  }
  if (o->point) {
    const double obj1 = pairwise_sum_f32(all_v, E), obj2 = pairwise_sum_f32(all_dt, N);
    double con1 = 0.0, con2 = 0.0;
    for (int e = 0; e < E; ++e) { const double v = fabs(all_s[e]); if (v > con1 || v != v) con1 = v; }
    for (int i = 0; i < N; ++i) { const double v = fabs(all_d[i]); if (v > con2 || v != v) con2 = v; }
    float* p = o->point + b * 4;
    p[0] = (float)f32(obj1 / (double)f->int_obj1);
    p[1] = (float)f32(obj2 / (double)f->int_obj2);
    p[2] = (float)con1; p[3] = (float)con2;
  }
}

/* Game_research04._game_modify for B environments.  a_geo / a_topo are clipped in place and stale_range (the max_up /
 * max_down left by the previous call, [B][N][2] float32) is read only.  Returns 0. */
int truss_oracle_step(const oracle_family* f, int B, const float* set_node, const float* set_element, float* a_geo,
                      float* a_topo, const uint8_t* coin, const float* stale_range, const oracle_out* o, int nthreads) {
  const int N = f->N, E = f->E;
  if (N > MAXN || E > MAXE || f->ndof > MAXDOF) return -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int bb = 0; bb < B; ++bb) {
    const size_t b = (size_t)bb;
    float* ag = a_geo + b * N * 2;
    float* at = a_topo + b * N * 3;
    clip01(ag, N * 2); clip01(at, N * 3);
    TS y[MAXN];
    int sec[MAXE];
    for (int i = 0; i < N; ++i) y[i] = S((double)set_node[(b * N + i) * 12 + 1]);
    for (int e = 0; e < E; ++e) sec[e] = (int)set_element[(b * E + e) * 21];
    for (int i = 0; i < N; ++i) {                                      /* action decode (:396-411) */
      const double row[2] = {(double)ag[2 * i], (double)ag[2 * i + 1]};
      const int adj = np_argmax(row, 2);
      const double v = row[adj];
      const TS amt = (v < 1) ? S(v) : W(1);                            /* min([1, a]) */
      const TS range = S((double)stale_range[(b * N + i) * 2 + adj]);
      const TS step = ts_mul(ts_mul(amt, range), W(0.25));
      y[i] = (adj == 0) ? ts_add(y[i], step) : ts_sub(y[i], step);
    }
    for (int i = 0; i < N; ++i) {                                      /* supports, round to cm (:413-416) */
      if (f->res[2 * i + 1] == 1) y[i] = W(0);
      if (!y[i].weak) y[i].v = round2_f32(y[i].v);                     /* weak values here are the supports' 0 */
    }
    for (int e = 0; e < E; ++e) {                                      /* section change (:419-431) */
      const int n0 = f->conn[2 * e], n1 = f->conn[2 * e + 1];
      double pv[3];
      for (int k = 0; k < 3; ++k) pv[k] = f32((double)at[3 * n0 + k] + (double)at[3 * n1 + k]);
      const int am = np_argmax(pv, 3);
      if (am == 0) sec[e] = sec[e] - 1 < 0 ? 0 : sec[e] - 1;
      else if (am == 1) sec[e] = sec[e] + 1 > NSEC - 1 ? NSEC - 1 : sec[e] + 1;
    }
    const TS ymin = W(f->y_min), ymax = W(f->y_max), dmin = W(f->d_min);
    for (int i = 0; i < N; ++i)                                        /* pass (i) (:434-440) */
      if (ts_lt(y[i], ymin)) {
        if (f->top[i] == 1) { y[i] = dmin; y[f->pair[i]] = ymin; } else y[i] = ymin;
      }
    for (int i = 0; i < N; ++i)                                        /* pass (ii) (:442-448) */
      if (ts_gt(y[i], ymax)) {
        if (f->top[i] == 1) y[i] = ymax; else { y[i] = ts_sub(ymax, dmin); y[f->pair[i]] = ymax; }
      }
    for (int i = 0; i < N; ++i)                                        /* pass (iii) (:450-455) */
      if (ts_lt(ts_abs(ts_sub(y[i], y[f->pair[i]])), dmin) && f->top[i] == 1) y[i] = ts_add(y[f->pair[i]], dmin);
    TS ys[MAXN];
    const int32_t* src = f->sym_src + (coin && coin[b] ? N : 0);       /* symmetry copy (:460-502 / :460-557) */
    for (int i = 0; i < N; ++i) ys[i] = y[src[i]];
    for (int p = 0; p < f->npairs; ++p) {                              /* mirrored elements take the smaller section */
      const int a = f->sym_pairs[2 * p], c2 = f->sym_pairs[2 * p + 1];
      const int m = sec[a] < sec[c2] ? sec[a] : sec[c2];
      sec[a] = m; sec[c2] = m;
    }
    TS up[MAXN], down[MAXN];
    move_range(f, ys, up, down);
    for (int i = 0; i < N; ++i) {
      if (o->y) o->y[b * N + i] = ys[i].v;
      if (o->weak) o->weak[b * N + i] = (uint8_t)ys[i].weak;
      if (o->move_range) { o->move_range[(b * N + i) * 2] = (float)f32(up[i].v); o->move_range[(b * N + i) * 2 + 1] = (float)f32(down[i].v); }
    }
    if (o->section) for (int e = 0; e < E; ++e) o->section[b * E + e] = sec[e];
    evaluate(f, ys, sec, o, b);
  }
  return 0;
}

/* Model.restore(); Model.gen_all() + objectives on explicit float64 heights (all weak) */
int truss_oracle_solve(const oracle_family* f, int B, const double* y64, const int32_t* section, const oracle_out* o,
                       int nthreads) {
  const int N = f->N, E = f->E;
  if (N > MAXN || E > MAXE || f->ndof > MAXDOF) return -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
#endif
  for (int bb = 0; bb < B; ++bb) {
    const size_t b = (size_t)bb;
    TS y[MAXN];
    int sec[MAXE];
    for (int i = 0; i < N; ++i) y[i] = W(y64[b * N + i]);
    for (int e = 0; e < E; ++e) sec[e] = section[b * E + e];
    evaluate(f, y, sec, o, b);
  }
  return 0;
}
