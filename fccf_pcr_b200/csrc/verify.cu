// verify.cu — quick_verify (FCCF.cpp:680-783) with the Ceres plane-to-plane refinement
// (LidarPlaneFactor FCCF.cpp:178-208, ceres_refine 210-249), score_range (1233-1251) and the
// top-k selection (1499-1544), one WARP per hypothesis.
//
// Plane association: lane a < F1 tests source plane a against all target planes (<= 16 x 16).
// Refinement: Ceres 1.14's trust-region Levenberg-Marquardt restated (DENSE_QR, Jacobi scaling,
// EigenQuaternionParameterization, <= 50 iterations; App. A.5 of SURVEY.md).  Lane l owns residual
// row l (pair l/2, residual l%2) and, for l < 6, the l-th damping row of the augmented system
// [J; sqrt(D)] y = [r; 0]; every "sum over rows" is one xor-butterfly over the warp (all lanes end
// with the same bits), the Householder QR runs column by column on those register rows.  FP64
// throughout — this is the FP64-pipe stage of the path.
//
// In the pipeline the refinement is LAZY: the ranking score is the pre-refinement one (Q16) and only
// the fine_verify_number best centres per type are read again, so quick_verify_kernel associates and
// scores every centre, rank_top_kernel ranks them, and refine_top_kernel runs the refinement for the
// <= 3 x 4 selected ones — instead of ~150 solves of up to 50 iterations per registration.  The
// stand-alone entry point fccf_quick_verify still refines every hypothesis it is given.
#include "fccf_dev.cuh"
#include "fccf_internal.h"
#include <cstdlib>
#include <vector>

namespace fccf {

__device__ __forceinline__ double bfly(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = v + __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void crossd(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dot3d(const double a[3], const double b[3]) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

// f(q,a) = a + w*uv + u x uv, uv = 2 (u x a); Jacobian wrt (x,y,z,w)
__device__ __forceinline__ void rot_with_jac(const double q[4], const double a[3], double f[3], double J[3][4], bool want) {
  const double* u = q; double w = q[3];
  double uv[3]; crossd(u, a, uv); uv[0] += uv[0]; uv[1] += uv[1]; uv[2] += uv[2];
  double c2[3]; crossd(u, uv, c2);
  #pragma unroll
  for (int i = 0; i < 3; i++) f[i] = (a[i] + w * uv[i]) + c2[i];
  if (!want) return;
  #pragma unroll
  for (int k = 0; k < 3; k++) {
    double e[3] = {0, 0, 0}; e[k] = 1.0;
    double Av[3]; crossd(e, a, Av); Av[0] *= 2; Av[1] *= 2; Av[2] *= 2;
    double t1[3], t2[3]; crossd(e, uv, t1); crossd(u, Av, t2);
    #pragma unroll
    for (int i = 0; i < 3; i++) J[i][k] = w * Av[i] + t1[i] + t2[i];
  }
  #pragma unroll
  for (int i = 0; i < 3; i++) J[i][3] = uv[i];
}

struct LmRow { double n1[3], p1[3], n2[3], p2[3], w; bool active; int odd; };

// residual of this lane's row and (optionally) its 6 local Jacobian entries; warp-uniform return
// fin_r (optional): all residuals finite, whatever the Jacobian does
__device__ __forceinline__ bool lm_eval(const LmRow& R, const double x[7], double& r, double J[6], bool want, bool* fin_r = nullptr) {
  bool fin = true, finr = true;
  r = 0.0;
  if (want) for (int c = 0; c < 6; c++) J[c] = 0.0;
  if (R.active) {
    const double* q = x; const double* t = x + 4;
    double n2r[3], p2r[3], Jn[3][4], Jp[3][4];
    rot_with_jac(q, R.n2, n2r, Jn, want);
    rot_with_jac(q, R.p2, p2r, Jp, want);
    #pragma unroll
    for (int i = 0; i < 3; i++) p2r[i] += t[i];
    double cv[3]; crossd(R.n1, n2r, cv);
    double nc = sqrt(dot3d(cv, cv));
    double d = dot3d(R.n1, R.p1) - dot3d(n2r, p2r);
    double sd = sqrt(d * d);
    r = R.odd ? R.w * sd : R.w * nc;
    fin = isfinite(r); finr = fin;
    if (want) {
      double Ja[7];
      // even rows: w * (cv . (n1 x dn)) / nc, odd rows: w * (d * dd) / sd — numerator and denominator are selected per
      // lane and ONE division runs for the whole warp (the same operations per lane as the two-sided branch)
      const double den = R.odd ? sd : nc;
      #pragma unroll
      for (int col = 0; col < 4; col++) {
        double dn[3] = {Jn[0][col], Jn[1][col], Jn[2][col]}, dp[3] = {Jp[0][col], Jp[1][col], Jp[2][col]};
        double dc[3]; crossd(R.n1, dn, dc);
        const double ne = dot3d(cv, dc);
        const double dd = -(dot3d(dn, p2r) + dot3d(n2r, dp));
        const double no = d * dd;
        Ja[col] = R.w * ((R.odd ? no : ne) / den);
      }
      #pragma unroll
      for (int col = 0; col < 3; col++) Ja[4 + col] = R.odd ? R.w * (d * (-n2r[col]) / sd) : 0.0;
      // EigenQuaternionParameterization::ComputeJacobian (4x3)
      double Pm[4][3] = {{q[3], q[2], -q[1]}, {-q[2], q[3], q[0]}, {q[1], -q[0], q[3]}, {-q[0], -q[1], -q[2]}};
      #pragma unroll
      for (int lc = 0; lc < 3; lc++) { double s = 0; for (int a = 0; a < 4; a++) s += Ja[a] * Pm[a][lc]; J[lc] = s; }
      #pragma unroll
      for (int lc = 0; lc < 3; lc++) J[3 + lc] = Ja[4 + lc];
      #pragma unroll
      for (int lc = 0; lc < 6; lc++) fin = fin && isfinite(J[lc]);
    }
  }
  if (fin_r) *fin_r = __all_sync(0xffffffffu, finr);
  return __all_sync(0xffffffffu, fin);
}
__device__ __forceinline__ void lm_plus(const double x[7], const double delta[6], double out[7]) {
  double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
  if (nd > 0.0) {
    double s = sin(nd) / nd;
    double dq[4] = {s * delta[0], s * delta[1], s * delta[2], cos(nd)};
    const double* b = x;
    out[3] = dq[3] * b[3] - dq[0] * b[0] - dq[1] * b[1] - dq[2] * b[2];
    out[0] = dq[3] * b[0] + dq[0] * b[3] + dq[1] * b[2] - dq[2] * b[1];
    out[1] = dq[3] * b[1] + dq[1] * b[3] + dq[2] * b[0] - dq[0] * b[2];
    out[2] = dq[3] * b[2] + dq[2] * b[3] + dq[0] * b[1] - dq[1] * b[0];
  } else { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; }
  #pragma unroll
  for (int i = 0; i < 3; i++) out[4 + i] = x[4 + i] + delta[3 + i];
}
// min || [A; B] y - [b; 0] || by Householder QR.  Lane l holds main row l (A, b) and, for l < 6,
// augmented row l (B = diag(lmd)).  Returns false on a zero column / non-finite solution.
// Per column ONE round of butterflies: the products of column k with the columns behind it and with the right-
// hand side (g_j = a_k . a_j, all independent) give the norm (g_k), the reflector's v.v = g_k + alpha^2 - 2 alpha a_kk
// and every v . a_j = g_j - alpha a_kj — the same reflections as the three dependent reductions per column of the
// textbook loop, with a third of the dependent shuffle depth.  (The sums associate differently from the oracle's:
// agreement is to rounding, inside the 0.01 degree / 1 mm bar by orders of magnitude — tests/lm_check.py.)
__device__ __forceinline__ bool lm_qr_solve(double A[6], double b, double B[6], int lane, double y[6]) {
  double bb = 0.0;   // rhs of the augmented row
  const bool aug = lane < 6;
  #pragma unroll
  for (int k = 0; k < 6; k++) {
    const double mk = (lane >= k) ? A[k] : 0.0;
    const double ak = aug ? B[k] : 0.0;
    double g[7];
    #pragma unroll
    for (int j = k; j < 6; j++) g[j] = mk * A[j] + ak * (aug ? B[j] : 0.0);
    g[6] = mk * b + ak * bb;
    #pragma unroll
    for (int o = 16; o; o >>= 1) {
      #pragma unroll
      for (int j = k; j < 7; j++) g[j] = g[j] + __shfl_xor_sync(0xffffffffu, g[j], o);
    }
    const double nrm = sqrt(g[k]);
    if (nrm == 0.0) return false;
    const double akk = __shfl_sync(0xffffffffu, A[k], k);
    const double alpha = (akk > 0) ? -nrm : nrm;
    const double vtv = (g[k] + alpha * alpha) - 2.0 * (alpha * akk);
    if (vtv == 0.0) return false;
    const double beta = 2.0 / vtv;
    const double vm = (lane == k) ? (akk - alpha) : mk;     // Householder vector entries of this lane's rows
    const double va = ak;
    #pragma unroll
    for (int j = k + 1; j < 6; j++) {
      const double akj = __shfl_sync(0xffffffffu, A[j], k);
      const double s = (g[j] - alpha * akj) * beta;
      A[j] -= s * vm;
      if (aug) B[j] -= s * va;
    }
    {
      const double bk = __shfl_sync(0xffffffffu, b, k);
      const double s = (g[6] - alpha * bk) * beta;
      b -= s * vm;
      if (aug) bb -= s * va;
    }
    if (lane == k) A[k] = alpha;
  }
  // back substitution: the six reciprocals of the diagonal first (independent divisions), then multiplications
  double amine = A[0];             // row k's diagonal entry lives in lane k: its reciprocal is computed there, once
  #pragma unroll
  for (int k = 1; k < 6; k++) if (lane == k) amine = A[k];
  const double dinv = 1.0 / amine;
  bool ok = true;
  #pragma unroll
  for (int k = 5; k >= 0; k--) {
    double s = b;
    #pragma unroll
    for (int j = k + 1; j < 6; j++) s -= A[j] * y[j];
    y[k] = __shfl_sync(0xffffffffu, s * dinv, k);
    ok = ok && isfinite(y[k]);
  }
  return ok;
}

// ceres::Solve for one hypothesis (whole warp).  x = (qx,qy,qz,qw,tx,ty,tz)
__device__ int lm_refine_warp(const LmRow& R, int lane, double x[7]) {
  const int max_iter = 50;
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16, min_diag = 1e-6, max_diag = 1e32;
  x[0] = 0; x[1] = 0; x[2] = 0; x[3] = 1; x[4] = 0; x[5] = 0; x[6] = 0;
  double r, J[6];
  double radius = 1e4, decrease_factor = 2.0; bool reuse_diag = false;
  double scale[6], diag[6], g[6];
  int iter = 0;
  if (!lm_eval(R, x, r, J, true)) return 0;
  double cost = 0.5 * bfly(r * r);
  #pragma unroll
  for (int c = 0; c < 6; c++) g[c] = bfly(J[c] * r);
  #pragma unroll
  for (int c = 0; c < 6; c++) scale[c] = 1.0 / (1.0 + sqrt(bfly(J[c] * J[c])));
  #pragma unroll
  for (int c = 0; c < 6; c++) J[c] *= scale[c];
  double gmax;
  {
    double ng[6], xp[7]; for (int c = 0; c < 6; c++) ng[c] = -g[c];
    lm_plus(x, ng, xp);
    gmax = 0; for (int i = 0; i < 7; i++) gmax = fmax(gmax, fabs(x[i] - xp[i]));
  }
  double x_norm = 0; for (int i = 0; i < 7; i++) x_norm += x[i] * x[i]; x_norm = sqrt(x_norm);
  int invalid = 0; bool step_successful = true;
  while (true) {
    if (iter >= max_iter) break;
    if (step_successful && gmax <= gradient_tolerance) break;
    if (radius < min_radius) break;
    iter++;
    step_successful = false;
    if (!reuse_diag) for (int c = 0; c < 6; c++) diag[c] = fmin(fmax(bfly(J[c] * J[c]), min_diag), max_diag);
    double Aq[6], Bq[6], step[6];
    // damping row c belongs to lane c: one division and one square root per lane, not six
    double dmine = diag[0];
    #pragma unroll
    for (int c = 1; c < 6; c++) if (lane == c) dmine = diag[c];
    const double bmine = sqrt(dmine / radius);
    #pragma unroll
    for (int c = 0; c < 6; c++) { Aq[c] = J[c]; Bq[c] = (lane == c) ? bmine : 0.0; }
    bool solved = lm_qr_solve(Aq, r, Bq, lane, step);
    reuse_diag = true;
    bool valid = false; double model_change = 0;
    if (solved) {
      #pragma unroll
      for (int c = 0; c < 6; c++) step[c] = -step[c];
      double mr = 0; for (int c = 0; c < 6; c++) mr += J[c] * step[c];
      model_change = -bfly(mr * (r + mr / 2.0));
      valid = (model_change > 0.0);
    }
    if (!valid) {
      invalid++;
      if (invalid >= 5) break;
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
      continue;
    }
    invalid = 0;
    double delta[6]; for (int c = 0; c < 6; c++) delta[c] = step[c] * scale[c];
    double xc[7]; lm_plus(x, delta, xc);
    // residuals AND Jacobian at the candidate in one evaluation: an accepted step (nearly all of them) needs both
    double rc, Jd[6];
    double cand_cost;
    bool cand_fin_r;
    const bool cand_fin = lm_eval(R, xc, rc, Jd, true, &cand_fin_r);
    if (cand_fin_r) cand_cost = 0.5 * bfly(rc * rc);
    else cand_cost = 1.7976931348623157e308;
    double sn = 0; for (int i = 0; i < 7; i++) sn += (x[i] - xc[i]) * (x[i] - xc[i]); sn = sqrt(sn);
    if (sn <= parameter_tolerance * (x_norm + parameter_tolerance)) break;
    double cost_change = cost - cand_cost;
    if (fabs(cost_change) <= function_tolerance * cost) break;
    double rel = cost_change / model_change;
    if (rel > min_relative_decrease) {
      #pragma unroll
      for (int i = 0; i < 7; i++) x[i] = xc[i];
      x_norm = 0; for (int i = 0; i < 7; i++) x_norm += x[i] * x[i]; x_norm = sqrt(x_norm);
      cost = cand_cost;
      if (!cand_fin) break;
      r = rc;
      #pragma unroll
      for (int c = 0; c < 6; c++) J[c] = Jd[c];
      #pragma unroll
      for (int c = 0; c < 6; c++) g[c] = bfly(J[c] * r);
      #pragma unroll
      for (int c = 0; c < 6; c++) J[c] *= scale[c];
      {
        double ng[6], xp[7]; for (int c = 0; c < 6; c++) ng[c] = -g[c];
        lm_plus(x, ng, xp);
        gmax = 0; for (int i = 0; i < 7; i++) gmax = fmax(gmax, fabs(x[i] - xp[i]));
      }
      step_successful = true;
      { const double u = 2.0 * rel - 1.0; radius = radius / fmax(1.0 / 3.0, 1.0 - u * u * u); }   // pow(u, 3)
      radius = fmin(max_radius, radius);
      decrease_factor = 2.0; reuse_diag = false;
    } else {
      radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diag = true;
    }
  }
  return iter;
}

// quick_verify for one hypothesis by one warp.  planes: stride 8 floats (c, n, size, -).
// REFINE=false: association and score only (T is left as given); iters_out gets -2 where a refinement
// would have run.
template <bool REFINE>
__device__ float quick_verify_warp(float T[16], const float* pl1, int F1, const float* pl2, int F2, float ang_cut, float dist_thr,
                                   float required, int lane, int* npair_out, int* pairs_out, int* iters_out) {
  // integer-truncating size sums (FCCF.cpp:693,707)
  int fs1 = 0, fs2 = 0;
  for (int k = 0; k < F1; k++) fs1 = (int)((float)fs1 + pl1[k * 8 + 6]);
  for (int k = 0; k < F2; k++) fs2 = (int)((float)fs2 + pl2[k * 8 + 6]);
  bool find = false; int best = 0; float best_imp = 0.f, best_score = 0.f;
  f3 p1 = mk3(0, 0, 0), n1 = mk3(0, 0, 0);
  if (lane < F1) {
    p1 = mk3(pl1[lane * 8], pl1[lane * 8 + 1], pl1[lane * 8 + 2]); n1 = mk3(pl1[lane * 8 + 3], pl1[lane * 8 + 4], pl1[lane * 8 + 5]);
    float size1 = pl1[lane * 8 + 6];
    double d1d = sum3d((double)n1.x * (double)p1.x, (double)n1.y * (double)p1.y, (double)n1.z * (double)p1.z);
    float d1 = (float)d1d;
    const double nn1 = normal_norm(n1.x, n1.y, n1.z);
    for (int b = 0; b < F2; b++) {
      f3 p2 = tf_se3(T, mk3(pl2[b * 8], pl2[b * 8 + 1], pl2[b * 8 + 2]));
      f3 n2 = tf_so3(T, mk3(pl2[b * 8 + 3], pl2[b * 8 + 4], pl2[b * 8 + 5]));
      // compute_normal_angel(...) < quick_verify_angel_threshold (FCCF.cpp:722-730) through its cosine cut
      bool ang_ok = angle_lt(normal_cos_n(n1.x, n1.y, n1.z, nn1, n2.x, n2.y, n2.z, normal_norm(n2.x, n2.y, n2.z)), ang_cut);
      float d2 = (float)sum3d((double)n2.x * (double)p2.x, (double)n2.y * (double)p2.y, (double)n2.z * (double)p2.z);
      float dist = (float)fabs((double)(d1 - d2));
      if (ang_ok && dist < dist_thr) {
        find = true;
        float size2 = pl2[b * 8 + 6];
        float mn = size1 < size2 ? size1 : size2, mx = size1 > size2 ? size1 : size2;
        float cs = mn / mx;
        float ci = (2 * mn) / (float)(fs1 + fs2);
        if (cs > best_score) { best_imp = ci; best_score = cs; best = b; }
      }
    }
  }
  unsigned fm = __ballot_sync(0xffffffffu, find);
  int np = __popc(fm);
  // row layout: lane l -> pair l/2
  int k = lane >> 1;
  int src = (k < np) ? (int)__fns(fm, 0, k + 1) : 0;
  f3 bp2 = mk3(0, 0, 0), bn2 = mk3(0, 0, 0);
  if (find) {
    bp2 = tf_se3(T, mk3(pl2[best * 8], pl2[best * 8 + 1], pl2[best * 8 + 2]));
    bn2 = tf_so3(T, mk3(pl2[best * 8 + 3], pl2[best * 8 + 4], pl2[best * 8 + 5]));
  }
  LmRow R;
  R.n1[0] = __shfl_sync(0xffffffffu, n1.x, src); R.n1[1] = __shfl_sync(0xffffffffu, n1.y, src); R.n1[2] = __shfl_sync(0xffffffffu, n1.z, src);
  R.p1[0] = __shfl_sync(0xffffffffu, p1.x, src); R.p1[1] = __shfl_sync(0xffffffffu, p1.y, src); R.p1[2] = __shfl_sync(0xffffffffu, p1.z, src);
  R.n2[0] = __shfl_sync(0xffffffffu, bn2.x, src); R.n2[1] = __shfl_sync(0xffffffffu, bn2.y, src); R.n2[2] = __shfl_sync(0xffffffffu, bn2.z, src);
  R.p2[0] = __shfl_sync(0xffffffffu, bp2.x, src); R.p2[1] = __shfl_sync(0xffffffffu, bp2.y, src); R.p2[2] = __shfl_sync(0xffffffffu, bp2.z, src);
  float wimp = __shfl_sync(0xffffffffu, best_imp, src);
  int bsel = __shfl_sync(0xffffffffu, best, src);
  R.w = (double)wimp; R.active = (k < np); R.odd = lane & 1;
  if (pairs_out && R.active && !(lane & 1)) { pairs_out[2 * k] = src; pairs_out[2 * k + 1] = bsel; }
  if (npair_out && lane == 0) *npair_out = np;
  int iters = -1;
  if (!REFINE && (float)np >= required) iters = -2;
  if (REFINE && (float)np >= required) {
    double x[7];
    iters = lm_refine_warp(R, lane, x);
    q4 q; q.w = (float)x[3]; q.x = (float)x[0]; q.y = (float)x[1]; q.z = (float)x[2];   // FCCF.cpp:231-236
    m3 Rm = quat_to_matrix(q);
    float N[16] = {Rm.m[0][0], Rm.m[0][1], Rm.m[0][2], (float)x[4], Rm.m[1][0], Rm.m[1][1], Rm.m[1][2], (float)x[5],
                   Rm.m[2][0], Rm.m[2][1], Rm.m[2][2], (float)x[6], 0.f, 0.f, 0.f, 1.f};
    float O[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {   // Eigen 4x4 packet product: sequential k
      float s = N[4 * i] * T[j];
      s = N[4 * i + 1] * T[4 + j] + s;
      s = N[4 * i + 2] * T[8 + j] + s;
      s = N[4 * i + 3] * T[12 + j] + s;
      O[4 * i + j] = s;
    }
    for (int i = 0; i < 16; i++) T[i] = O[i];
  }
  if (iters_out && lane == 0) *iters_out = iters;
  float score = 0.f;
  for (int kk = 0; kk < np; kk++) score = score + __shfl_sync(0xffffffffu, wimp, 2 * kk);
  return score;
}

struct QvArgs {
  PipeState* st;
  const float* centre; float* qv_T; float* qv_score; int* qv_npair; int* qv_pairs; int* qv_iters;
  int* rank_perm; float* top_T; float* top_s1; int* top_centre;
  int topk;
  float ang_cut, dist_thr, required, fine_number;   // ang_cut: cosine cut of quick_verify_angel_threshold (strict <)
};

// Plane association and score of every cluster centre (FCCF.cpp:1468-1492 without the refinement): one
// warp per centre.  The score that ranks the centres is computed from the pairs found BEFORE the
// refinement (Q16), and only the fine_verify_number best centres per type are ever read again
// (FCCF.cpp:1499-1544), so the Levenberg-Marquardt refinement — by far the most expensive part of
// quick_verify — runs only for those (refine_top_kernel), with bit-identical results for everything the
// path consumes.
__global__ void __launch_bounds__(128) quick_verify_kernel(const QvArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const QvArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int ty = wid / FCCF_MAXCENTRE, ci = wid - ty * FCCF_MAXCENTRE;
  if (ty >= 3 || ci >= st->n_centre[ty]) return;
  const float* c = A.centre + (size_t)wid * 8;
  q4 q; q.w = c[0]; q.x = c[1]; q.y = c[2]; q.z = c[3];
  m3 Rm = quat_to_matrix(q);   // FCCF.cpp:1470-1489
  float T[16] = {Rm.m[0][0], Rm.m[0][1], Rm.m[0][2], c[4], Rm.m[1][0], Rm.m[1][1], Rm.m[1][2], c[5], Rm.m[2][0], Rm.m[2][1], Rm.m[2][2], c[6], 0.f, 0.f, 0.f, 1.f};
  float s = quick_verify_warp<false>(T, &st->ft[0].plane[0][0], st->ft[0].F, &st->ft[1].plane[0][0], st->ft[1].F, A.ang_cut, A.dist_thr, A.required, lane,
                                     A.qv_npair + wid, A.qv_pairs + (size_t)wid * 32, A.qv_iters + wid);
  if (lane < 16) A.qv_T[(size_t)wid * 16 + lane] = T[lane];
  if (lane == 0) A.qv_score[wid] = s;
}

// score_range + top-k (FCCF.cpp:1494-1544): one CTA per type ranks the centres
__global__ void __launch_bounds__(256) rank_top_kernel(const QvArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const QvArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int t = threadIdx.x, ty = blockIdx.x;
  __shared__ float s_key[FCCF_MAXCENTRE];
  __shared__ int s_perm[FCCF_MAXCENTRE];
  __shared__ unsigned long long s_sort[40];
  const int C = st->n_centre[ty];
  int has_nan = 0;
  for (int k = t; k < C; k += 256) { float v = A.qv_score[ty * FCCF_MAXCENTRE + k]; s_key[k] = v; s_perm[k] = k; has_nan |= (v != v); }
  has_nan = __syncthreads_or(has_nan);
  int amax = (int)A.fine_number;
  if (amax > A.topk) amax = A.topk;
  int nt = C < amax ? C : amax;
  // Only the first nt positions of the exchange sort are ever read (FCCF.cpp:1499-1544): pass i of the sort makes
  // position i final, so a few selected centres need a few literal passes by one warp, not the whole sort.
  if (nt <= 16 || has_nan) { if (t < 32) warp_exchange_sort(s_key, s_perm, C, [](float a, float b) { return a < b; }, has_nan ? 0x7fffffff : nt); }   // NaN keys have no total order: literal emulation
  else block_exchange_sort(s_key, s_perm, C, s_sort);
  __syncthreads();
  for (int k = t; k < C; k += 256) A.rank_perm[ty * FCCF_MAXCENTRE + k] = s_perm[k];   // (complete only when the whole sort ran)
  for (int k = t; k < nt; k += 256) { A.top_s1[ty * A.topk + k] = s_key[k]; A.top_centre[ty * A.topk + k] = s_perm[k]; }
  if (t == 0) st->n_top[ty] = nt;
}

// quick_verify with the Ceres refinement for the selected centres: one warp per (type, rank)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) refine_top_kernel(const QvArgs* __restrict__ AB) {
  FCCF_PDL_ENTER();
  const QvArgs& A = AB[blockIdx.z];
  PipeState* st = A.st;
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int ty = slot / A.topk, k = slot - ty * A.topk;
  if (ty >= 3 || k >= st->n_top[ty]) return;
  const int wid = ty * FCCF_MAXCENTRE + A.top_centre[slot];
  const float* c = A.centre + (size_t)wid * 8;
  q4 q; q.w = c[0]; q.x = c[1]; q.y = c[2]; q.z = c[3];
  m3 Rm = quat_to_matrix(q);
  float T[16] = {Rm.m[0][0], Rm.m[0][1], Rm.m[0][2], c[4], Rm.m[1][0], Rm.m[1][1], Rm.m[1][2], c[5], Rm.m[2][0], Rm.m[2][1], Rm.m[2][2], c[6], 0.f, 0.f, 0.f, 1.f};
  quick_verify_warp<true>(T, &st->ft[0].plane[0][0], st->ft[0].F, &st->ft[1].plane[0][0], st->ft[1].F, A.ang_cut, A.dist_thr, A.required, lane,
                          nullptr, nullptr, A.qv_iters + wid);
  if (lane < 16) { A.top_T[(size_t)slot * 16 + lane] = T[lane]; A.qv_T[(size_t)wid * 16 + lane] = T[lane]; }
}

void launch_quick_verify(cudaStream_t s, const Batch& b, uint64_t* launches) {
  const int G = b.G;
  std::vector<QvArgs> As(G);
  for (int g = 0; g < G; g++) {
    const Work& w = b.w[g]; const HypWS& h = w.h;
    QvArgs& A = As[g];
    A.st = w.st; A.centre = h.centre; A.qv_T = h.qv_T; A.qv_score = h.qv_score; A.qv_npair = h.qv_npair; A.qv_pairs = h.qv_pairs; A.qv_iters = h.qv_iters;
    A.rank_perm = h.rank_perm; A.top_T = h.top_T; A.top_s1 = h.top_s1; A.top_centre = h.top_centre;
    A.ang_cut = b.cuts.qv_lt; A.dist_thr = b.p.quick_verify_distance_threshold; A.required = b.p.required_optimize_plane; A.fine_number = b.p.fine_verify_number; A.topk = fccf_topk(b.p);
  }
  const QvArgs* dA = b.tab->put(As.data(), G);
  static int occ = -1;
  if (occ < 0) { const char* e = getenv("FCCF_QV_OCC"); occ = e ? atoi(e) : 2; }
  klaunch(quick_verify_kernel, dim3(dim3((3 * FCCF_MAXCENTRE + 3) / 4, 1, G)), dim3(128), 0, s, dA);
  klaunch(rank_top_kernel, dim3(dim3(3, 1, G)), dim3(256), 0, s, dA);
  dim3 grid((3 * fccf_topk(b.p) + 3) / 4, 1, G);
  if (occ <= 2) klaunch(refine_top_kernel<2>, dim3(grid), dim3(128), 0, s, dA);
  else if (occ == 3) klaunch(refine_top_kernel<3>, dim3(grid), dim3(128), 0, s, dA);
  else klaunch(refine_top_kernel<4>, dim3(grid), dim3(128), 0, s, dA);
  if (launches) *launches += 3;
}

// stand-alone: n hypotheses (row-major 4x4, updated in place) against two plane tables (F x 8)
struct QvListArgs { float* T; int n; const float* pl1; int f1; const float* pl2; int f2; float* score; int* npair; int* pairs; int* iters; float ang_cut, dist_thr, required; };
__global__ void __launch_bounds__(128) quick_verify_list_kernel(const __grid_constant__ QvListArgs A) {
  FCCF_PDL_ENTER();
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (wid >= A.n) return;
  float T[16];
  for (int i = 0; i < 16; i++) T[i] = A.T[(size_t)wid * 16 + i];
  float s = quick_verify_warp<true>(T, A.pl1, A.f1, A.pl2, A.f2, A.ang_cut, A.dist_thr, A.required, lane, A.npair ? A.npair + wid : nullptr,
                              A.pairs ? A.pairs + (size_t)wid * 32 : nullptr, A.iters ? A.iters + wid : nullptr);
  if (lane < 16) A.T[(size_t)wid * 16 + lane] = T[lane];
  if (lane == 0) A.score[wid] = s;
}
void launch_quick_verify_list(cudaStream_t s, const fccf_params& p, float* d_T16, int n, const float* d_planes1, int f1,
                              const float* d_planes2, int f2, float* d_score, int* d_npair, int* d_pairs, int* d_iters, uint64_t* launches) {
  QvListArgs A;
  A.T = d_T16; A.n = n; A.pl1 = d_planes1; A.f1 = f1; A.pl2 = d_planes2; A.f2 = f2; A.score = d_score; A.npair = d_npair; A.pairs = d_pairs; A.iters = d_iters;
  A.ang_cut = angle_cut(p.quick_verify_angel_threshold, true); A.dist_thr = p.quick_verify_distance_threshold; A.required = p.required_optimize_plane;
  if (n <= 0) return;
  klaunch(quick_verify_list_kernel, dim3((n + 3) / 4), dim3(128), 0, s, A);
  if (launches) *launches += 1;
}

}  // namespace fccf
