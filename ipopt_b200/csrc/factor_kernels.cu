// Numeric factorisation kernels of the B200 supernodal multifrontal LDL^T (sm_100a).
//
// What the reference delegates to the vendor library's numeric phase (MUMPS job=2,
// reference src/Algorithm/LinearSolvers/IpMumpsSolverInterface.cpp:448-541): per front
// assemble original entries + children contribution blocks (extend-add), eliminate the
// fully-summed block with 1x1/2x2 Bunch-Kaufman pivots (threshold-tested against the whole
// front column), form the Schur complement, hand the contribution block to the parent.
//
// Two execution classes:
//   * fronts of order <= smem_front_max: ONE CTA does gather + extend-add + pivoted
//     factorisation + Schur update entirely in shared memory and writes L / CB once;
//   * larger fronts: global-memory blocked right-looking algorithm (32-wide panels) made of
//     batched kernels (diag-block factor, panel TRSM, trailing update, Schur GEMM).
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.cuh"

namespace b200 {

#define BK_ALPHA 0.6403882032022076  // (1+sqrt(17))/8
#define NB 32                        // panel width of the big-front path

// --------------------------------------------------------------------------------------------
// small utility kernels
// --------------------------------------------------------------------------------------------
__global__ void k_sum_dups(long long nu, const long long* __restrict__ useg_ptr,
                           const int* __restrict__ useg_src, const double* __restrict__ vals,
                           double* __restrict__ uval) {
  long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (u >= nu) return;
  double acc = 0.0;
  for (long long q = useg_ptr[u]; q < useg_ptr[u + 1]; ++q) acc += vals[useg_src[q]];
  uval[u] = acc;
}

// rmax[i] = max_j |a_ij| s_i s_j  (bit pattern of non-negative doubles orders like integers)
__global__ void k_rowmax(long long nu, const int* __restrict__ u_row, const int* __restrict__ u_col,
                         const double* __restrict__ uval, const double* __restrict__ scale,
                         unsigned long long* __restrict__ rmax) {
  long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (u >= nu) return;
  int i = u_row[u], j = u_col[u];
  double a = fabs(uval[u]) * scale[i] * scale[j];
  if (!(a > 0.0)) return;  // also skips NaN
  unsigned long long b = (unsigned long long)__double_as_longlong(a);
  atomicMax(rmax + i, b);
  if (i != j) atomicMax(rmax + j, b);
}

__global__ void k_scale_update(int n, double* __restrict__ scale, unsigned long long* __restrict__ rmax) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double m = __longlong_as_double((long long)rmax[i]);
  rmax[i] = 0ull;
  if (m > 0.0 && isfinite(m)) {
    // power-of-two factor closest to 1/sqrt(m): scaling stays exact in floating point
    int e;
    frexp(m, &e);  // m = f * 2^e, f in [0.5,1)
    scale[i] *= ldexp(1.0, -((e) / 2));
  }
}

__global__ void k_apply_scale(long long nu, const int* __restrict__ u_row, const int* __restrict__ u_col,
                              const double* __restrict__ scale, double* __restrict__ uval) {
  long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (u >= nu) return;
  uval[u] *= scale[u_row[u]] * scale[u_col[u]];
}

__global__ void k_fill(int n, double* x, double v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

// --------------------------------------------------------------------------------------------
// pivoted LDL^T of a dense symmetric front held in shared memory (full square storage).
//   F   : f x f, leading dimension ld, both triangles valid on entry
//   k   : number of fully-summed (pivot) columns, candidates are rows/cols [0,k)
//   lp  : k ints, out: pivot position t holds original local column lp[t]
//   pt  : k ints, out: pivot type (1, 2, 3)
//   cv1, cv2 : f doubles scratch;  sh : >= 4 ints scratch
//   dinv/doff : global arrays already offset to this front's first pivot
// On exit columns [0,k) of F hold L (strictly lower part; F[t+1,t] of a 2x2 pivot holds the
// off-diagonal of D and must be written out as 0) and F[k:, k:] holds the Schur complement.
// All threads of the CTA must call this; blockDim.x is a multiple of 32.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_argmax(double& v, int& idx) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi >= 0 && (idx < 0 || oi < idx))) { v = ov; idx = oi; }
  }
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// The pivoted front factorisation runs either on a whole CTA (WARP=false, __syncthreads) or on a single
// warp (WARP=true, __syncwarp; several fronts per CTA).
template <bool WARP>
__device__ __forceinline__ void gsync() { if (WARP) __syncwarp(); else __syncthreads(); }

template <bool WARP>
__device__ void swap_sym(double* F, int ld, int f, int p, int q) {
  // symmetric interchange of rows/cols p and q of the full square; caller syncs before.
  const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x, nt = WARP ? 32 : blockDim.x;
  for (int t = tid; t < f; t += nt) {
    double a = F[p + t * ld], b = F[q + t * ld];
    F[p + t * ld] = b; F[q + t * ld] = a;
  }
  gsync<WARP>();
  for (int t = tid; t < f; t += nt) {
    double a = F[t + p * ld], b = F[t + q * ld];
    F[t + p * ld] = b; F[t + q * ld] = a;
  }
  gsync<WARP>();
}

template <bool WARP>
__device__ void factor_front_smem(double* F, int ld, int f, int k, int* lp, int* pt, double* cv1,
                                  double* cv2, volatile int* sh, double u, double tiny,
                                  double* dinv, double* doff, int* ptype_g, int* counters) {
  const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x, nt = WARP ? 32 : blockDim.x;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  for (int t = tid; t < k; t += nt) lp[t] = t;
  if (tid == 0) sh[2] = 0;
  int c_neg = 0, c_forced = 0, c_tiny = 0, c_2x2 = 0;  // only meaningful on thread 0
  int j = 0, kend = k, progress = 0;
  bool forced = false;
  gsync<WARP>();
  while (j < k) {
    if (j == kend) {  // no candidate left in this pass
      if (progress > 0) { kend = k; progress = 0; }
      else { forced = true; kend = k; }
    }
    // ---------------- pivot decision by warp 0 ----------------
    if (warp == 0) {
      int type = 0, r = -1;
      if (forced) {
        // every remaining candidate failed the threshold test: force 1x1 pivots.  If the whole remaining
        // column is rounding noise (scaled matrix, entries O(1)) the matrix is numerically singular.
        double cm = 0.0;
        for (int i = j + lane; i < f; i += 32) cm = fmax(cm, fabs(F[i + j * ld]));
        cm = warp_max(cm);
        type = 1; r = j;
        if (lane == 0) sh[2] = !(cm > 1e-12);
      } else {
        double lam = 0.0, gam = 0.0;
        int ridx = -1;
        for (int i = j + 1 + lane; i < f; i += 32) {
          double v = fabs(F[i + j * ld]);
          if (i < kend) { if (v > lam) { lam = v; ridx = i; } }
          else gam = fmax(gam, v);
        }
        warp_argmax(lam, ridx);
        gam = warp_max(gam);
        double ajj = fabs(F[j + j * ld]);
        bool ok1 = (ajj > tiny) && (ajj >= u * fmax(lam, gam));
        if (lam == 0.0 || ridx < 0) { if (ok1) { type = 1; r = j; } }
        else if (ok1 && ajj >= BK_ALPHA * lam) { type = 1; r = j; }
        else {
          r = ridx;
          double sig = 0.0, gamr = 0.0, cj = 0.0, cr = 0.0;
          for (int m = j + lane; m < f; m += 32) {
            if (m == r) continue;
            double v = (m < r) ? fabs(F[r + m * ld]) : fabs(F[m + r * ld]);
            if (m < kend) sig = fmax(sig, v); else gamr = fmax(gamr, v);
            if (m != j) { cr = fmax(cr, v); cj = fmax(cj, fabs(F[m + j * ld])); }
          }
          sig = warp_max(sig); gamr = warp_max(gamr); cj = warp_max(cj); cr = warp_max(cr);
          double arr = fabs(F[r + r * ld]);
          if (ok1 && ajj * sig >= BK_ALPHA * lam * lam) { type = 1; r = j; }
          else if (arr > tiny && arr >= BK_ALPHA * sig && arr >= u * fmax(sig, gamr)) { type = 2; }
          else {
            double a = F[j + j * ld], b = F[r + j * ld], c = F[r + r * ld];
            double det = a * c - b * b, adet = fabs(det);
            if (lam > tiny && adet > 0.0 && isfinite(adet) &&
                (fabs(c) * cj + fabs(b) * cr) * u <= adet && (fabs(a) * cr + fabs(b) * cj) * u <= adet)
              type = 3;
          }
        }
      }
      if (lane == 0) { sh[0] = type; sh[1] = r; }
    }
    gsync<WARP>();
    const int type = sh[0], r = sh[1], noise = sh[2];
    gsync<WARP>();  // everyone has read sh before it is rewritten
    if (type == 0) {  // reject for now: park column j at the end of the candidate range
      if (j != kend - 1) {
        swap_sym<WARP>(F, ld, f, j, kend - 1);
        if (tid == 0) { int t = lp[j]; lp[j] = lp[kend - 1]; lp[kend - 1] = t; }
      }
      --kend;
      continue;
    }
    if (type == 2 && r != j) {
      swap_sym<WARP>(F, ld, f, j, r);
      if (tid == 0) { int t = lp[j]; lp[j] = lp[r]; lp[r] = t; }
    }
    if (type == 3 && r != j + 1) {
      swap_sym<WARP>(F, ld, f, j + 1, r);
      if (tid == 0) { int t = lp[j + 1]; lp[j + 1] = lp[r]; lp[r] = t; }
    }
    if (type != 3) {
      // ---------------- 1x1 pivot at j ----------------
      double d = F[j + j * ld];
      if (forced) {
        if (noise || !(fabs(d) > tiny)) {  // zero (or NaN) pivot: static perturbation, reported as SINGULAR
          d = (d < 0.0) ? -1.5e-8 : 1.5e-8;
          if (tid == 0) ++c_tiny;
        } else if (tid == 0) ++c_forced;
      }
      const double dv = d;
      gsync<WARP>();
      for (int i = j + 1 + tid; i < f; i += nt) {
        double c = F[i + j * ld];
        cv1[i] = c;
        F[i + j * ld] = c / dv;
      }
      if (tid == 0) {
        pt[j] = 1; dinv[j] = 1.0 / dv; doff[j] = 0.0; ptype_g[j] = 1;
        if (dv < 0.0) ++c_neg;
      }
      gsync<WARP>();
      for (int m = j + 1 + warp; m < f; m += nwarp) {
        const double cm = cv1[m];
        if (cm != 0.0)
          for (int i = j + 1 + lane; i < f; i += 32) F[i + m * ld] -= F[i + j * ld] * cm;
      }
      gsync<WARP>();
      j += 1;
    } else {
      // ---------------- 2x2 pivot at (j, j+1) ----------------
      const double a = F[j + j * ld], b = F[j + 1 + j * ld], c = F[j + 1 + (j + 1) * ld];
      const double det = a * c - b * b;
      gsync<WARP>();
      for (int i = j + 2 + tid; i < f; i += nt) {
        double c1 = F[i + j * ld], c2 = F[i + (j + 1) * ld];
        cv1[i] = c1; cv2[i] = c2;
        F[i + j * ld] = (c * c1 - b * c2) / det;
        F[i + (j + 1) * ld] = (a * c2 - b * c1) / det;
      }
      if (tid == 0) {
        pt[j] = 2; pt[j + 1] = 3; ptype_g[j] = 2; ptype_g[j + 1] = 3;
        dinv[j] = c / det; dinv[j + 1] = a / det; doff[j] = -b / det; doff[j + 1] = 0.0;
        ++c_2x2;
        if (det < 0.0) c_neg += 1; else if (a < 0.0) c_neg += 2;
      }
      gsync<WARP>();
      for (int m = j + 2 + warp; m < f; m += nwarp) {
        const double m1 = cv1[m], m2 = cv2[m];
        if (m1 != 0.0 || m2 != 0.0)
          for (int i = j + 2 + lane; i < f; i += 32)
            F[i + m * ld] -= F[i + j * ld] * m1 + F[i + (j + 1) * ld] * m2;
      }
      gsync<WARP>();
      j += 2;
    }
    ++progress;
  }
  if (tid == 0) {
    if (c_neg) atomicAdd(counters + CNT_NEG, c_neg);
    if (c_forced) atomicAdd(counters + CNT_FORCED, c_forced);
    if (c_tiny) atomicAdd(counters + CNT_TINY, c_tiny);
    if (c_2x2) atomicAdd(counters + CNT_2X2, c_2x2);
  }
  gsync<WARP>();
}

// --------------------------------------------------------------------------------------------
// Class S/M: one CTA per front, everything in shared memory.
// smem: F[ld*f] | cv1[f] | cv2[f] | lp[k] | pt[k] | sh[8]
// --------------------------------------------------------------------------------------------
template <bool WARP>
__global__ void k_front_smem(DevSym S, DevNum N, const int* __restrict__ front_list, int nfronts,
                             int smem_per_group /*bytes, WARP mode only*/) {
  extern __shared__ double smem[];
  const int group = WARP ? (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) : blockIdx.x;
  if (group >= nfronts) return;
  const int s = front_list[group];
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int f = k + r, ld = f | 1;
  double* F = WARP ? (double*)((char*)smem + (size_t)(threadIdx.x >> 5) * smem_per_group) : smem;
  double* cv1 = F + (size_t)ld * f;
  double* cv2 = cv1 + f;
  int* lp = (int*)(cv2 + f);
  int* pt = lp + k;
  int* sh = pt + k;
  const int tid = WARP ? (threadIdx.x & 31) : threadIdx.x, nt = WARP ? 32 : blockDim.x;

  for (int t = tid; t < ld * f; t += nt) F[t] = 0.0;
  gsync<WARP>();
  // original entries (unique -> no write conflicts)
  for (long long uu = S.uent_ptr[s] + tid; uu < S.uent_ptr[s + 1]; uu += nt) {
    unsigned d = S.u_dst[uu];
    int lr = d & 0xffffu, lc = d >> 16;
    double v = N.uval[uu];
    F[lr + lc * ld] = v;
    F[lc + lr * ld] = v;
  }
  gsync<WARP>();
  // extend-add of the children contribution blocks, one child at a time (deterministic)
  for (int q = S.child_ptr[s]; q < S.child_ptr[s + 1]; ++q) {
    const int c = S.child_idx[q];
    const int rc = (int)(S.rows_ptr[c + 1] - S.rows_ptr[c]);
    const double* __restrict__ cb = N.CB + S.cb_off[c];
    const int* __restrict__ rl = S.rel + S.rows_ptr[c];
    for (int jj = tid >> 5; jj < rc; jj += (nt >> 5)) {
      const int lj = rl[jj];
      for (int ii = jj + (tid & 31); ii < rc; ii += 32) {
        const int li = rl[ii];
        const double v = cb[ii + (size_t)jj * rc];
        F[li + lj * ld] += v;
        if (li != lj) F[lj + li * ld] += v;
      }
    }
    gsync<WARP>();
  }

  factor_front_smem<WARP>(F, ld, f, k, lp, pt, cv1, cv2, sh, N.u, N.tiny, N.dinv + c0, N.doff + c0,
                          N.ptype + c0, N.counters);

  // write L panel (f x k, ld = f), unit diagonal, zero strictly-upper part of the pivot block
  double* __restrict__ P = N.L + S.L_off[s];
  for (int t = tid >> 5; t < k; t += (nt >> 5)) {
    const int is2 = (pt[t] == 2);
    for (int i = tid & 31; i < f; i += 32) {
      double v;
      if (i < t) v = 0.0;
      else if (i == t) v = 1.0;
      else if (is2 && i == t + 1) v = 0.0;
      else v = F[i + t * ld];
      P[i + (size_t)t * f] = v;
    }
  }
  for (int t = tid; t < k; t += nt) N.lperm[c0 + t] = lp[t];
  // contribution block (lower part)
  double* __restrict__ cbo = N.CB + S.cb_off[s];
  for (int m = tid >> 5; m < r; m += (nt >> 5))
    for (int i = m + (tid & 31); i < r; i += 32) cbo[i + (size_t)m * r] = F[(k + i) + (k + m) * ld];
}

// --------------------------------------------------------------------------------------------
// Class L (big fronts), global-memory blocked path.
// --------------------------------------------------------------------------------------------
__global__ void k_big_zero(DevSym S, DevNum N, const int* __restrict__ front_list) {
  const int s = front_list[blockIdx.y];
  const long long k = S.sn_start[s + 1] - S.sn_start[s];
  const long long r = S.rows_ptr[s + 1] - S.rows_ptr[s];
  const long long f = k + r, nl = f * k, nc = r * r;
  double* P = N.L + S.L_off[s];
  double* C = N.CB + S.cb_off[s];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < nl + nc;
       t += (long long)gridDim.x * blockDim.x) {
    if (t < nl) P[t] = 0.0; else C[t - nl] = 0.0;
  }
}

__global__ void k_big_assemble(DevSym S, DevNum N, const int* __restrict__ front_list) {
  const int s = front_list[blockIdx.y];
  double* P = N.L + S.L_off[s];
  for (long long uu = S.uent_ptr[s] + blockIdx.x * (long long)blockDim.x + threadIdx.x;
       uu < S.uent_ptr[s + 1]; uu += (long long)gridDim.x * blockDim.x)
    P[S.u_dst64[uu]] = N.uval[uu];
}

// extend-add of the q-th child of every big front in the list (one pass per child rank =>
// no two CTAs touch the same parent entry; deterministic, no atomics)
__global__ void k_big_extend_add(DevSym S, DevNum N, const int* __restrict__ front_list, int q) {
  const int p = front_list[blockIdx.y];
  const int nch = S.child_ptr[p + 1] - S.child_ptr[p];
  if (q >= nch) return;
  const int c = S.child_idx[S.child_ptr[p] + q];
  const int rc = (int)(S.rows_ptr[c + 1] - S.rows_ptr[c]);
  const int kp = S.sn_start[p + 1] - S.sn_start[p];
  const int rp = (int)(S.rows_ptr[p + 1] - S.rows_ptr[p]);
  const int fp = kp + rp;
  const double* __restrict__ cb = N.CB + S.cb_off[c];
  const int* __restrict__ rl = S.rel + S.rows_ptr[c];
  double* P = N.L + S.L_off[p];
  double* C = N.CB + S.cb_off[p];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int jj = blockIdx.x * nwarp + warp; jj < rc; jj += gridDim.x * nwarp) {
    const int lj = rl[jj];
    for (int ii = jj + lane; ii < rc; ii += 32) {
      const int li = rl[ii];
      const double v = cb[ii + (size_t)jj * rc];
      if (lj < kp) P[li + (size_t)lj * fp] += v;
      else C[(li - kp) + (size_t)(lj - kp) * rp] += v;
    }
  }
}

// factor the NB x NB diagonal block at panel offset jb (pivoting restricted to the block).
// ONE WARP per front (warp-synchronous, no block barriers).
__global__ void __launch_bounds__(32) k_big_diag(DevSym S, DevNum N, const int* __restrict__ front_list, int jb) {
  __shared__ double B[33 * NB];
  __shared__ double cv1[NB], cv2[NB];
  __shared__ int lp[NB], pt[NB], sh[8];
  const int s = front_list[blockIdx.x];
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  if (jb >= k) return;
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int nb = min(NB, k - jb), ld = 33;
  double* P = N.L + S.L_off[s];
  const int lane = threadIdx.x;
  for (int j = 0; j < nb; ++j)
    if (lane >= j && lane < nb) {
      double v = P[(jb + lane) + (size_t)(jb + j) * f];
      B[lane + j * ld] = v;
      B[j + lane * ld] = v;
    }
  __syncwarp();
  factor_front_smem<true>(B, ld, nb, nb, lp, pt, cv1, cv2, sh, N.u, N.tiny, N.dinv + c0 + jb,
                          N.doff + c0 + jb, N.ptype + c0 + jb, N.counters);
  for (int j = 0; j < nb; ++j)
    if (lane < nb) {
      double v;
      if (lane < j) v = 0.0;
      else if (lane == j) v = 1.0;
      else if (pt[j] == 2 && lane == j + 1) v = 0.0;
      else v = B[lane + j * ld];
      P[(jb + lane) + (size_t)(jb + j) * f] = v;
    }
  if (lane < nb) {
    N.bperm[c0 + jb + lane] = lp[lane];
    N.lperm[c0 + jb + lane] = jb + lp[lane];
  }
}

// rows below the diagonal block: W = A_perm * L_bb^-T (= L*D), L = W * D^-1.
// CTAs with blockIdx.x >= nrowblk apply the block's row interchanges to the L columns on the left.
__global__ void __launch_bounds__(128) k_big_trsm(DevSym S, DevNum N, const int* __restrict__ front_list, int jb,
                                                  int nrowblk) {
  __shared__ double Lb[33 * NB];
  __shared__ double di[NB], dof[NB];
  __shared__ int pty[NB], bp[NB];
  const int s = front_list[blockIdx.y];
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  if (jb >= k) return;
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int nb = min(NB, k - jb);
  const int row0 = jb + nb;
  double* P = N.L + S.L_off[s];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x >= nrowblk) {
    // ---- left part: rows jb..jb+nb of columns [0, jb) get the block permutation ----
    const int c = ((int)blockIdx.x - nrowblk) * blockDim.x + tid;
    if (tid < nb) bp[tid] = N.bperm[c0 + jb + tid];
    __syncthreads();
    if (c >= jb) return;
    double tmp[NB];
    double* col = P + (size_t)c * f + jb;
#pragma unroll
    for (int t = 0; t < NB; ++t) tmp[t] = (t < nb) ? col[t] : 0.0;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      if (t < nb) {
        const int src = bp[t];
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < NB; ++q) v = (q == src) ? tmp[q] : v;
        col[t] = v;
      }
    }
    return;
  }
  if ((long long)blockIdx.x * blockDim.x >= f - row0) return;
  double* Wp = N.W + S.L_off[s];
  for (int t = tid; t < nb * nb; t += blockDim.x) {
    int i = t % nb, j = t / nb;
    Lb[i + j * 33] = P[(jb + i) + (size_t)(jb + j) * f];
  }
  if (tid < nb) {
    di[tid] = N.dinv[c0 + jb + tid]; dof[tid] = N.doff[c0 + jb + tid];
    pty[tid] = N.ptype[c0 + jb + tid]; bp[tid] = N.bperm[c0 + jb + tid];
  }
  __syncthreads();
  const int i = row0 + blockIdx.x * blockDim.x + tid;
  if (i >= f) return;
  double x[NB];
#pragma unroll
  for (int t = 0; t < NB; ++t) x[t] = (t < nb) ? P[i + (size_t)(jb + bp[t]) * f] : 0.0;
#pragma unroll
  for (int t = 0; t < NB; ++t) {
    if (t < nb) {
      double acc = x[t];
#pragma unroll
      for (int q = 0; q < NB; ++q)
        if (q < t) acc -= x[q] * Lb[t + q * 33];
      x[t] = acc;
    }
  }
  const double lim = 1.0 / N.u;
  int bad = 0;
#pragma unroll
  for (int t = 0; t < NB; ++t) {
    if (t < nb) {
      double l;
      const int ty = pty[t];
      if (ty == 1) l = x[t] * di[t];
      else if (ty == 2) l = x[t] * di[t] + x[(t + 1 < NB) ? t + 1 : t] * dof[t];
      else l = x[(t > 0) ? t - 1 : 0] * dof[(t > 0) ? t - 1 : 0] + x[t] * di[t];
      Wp[i + (size_t)(jb + t) * f] = x[t];
      P[i + (size_t)(jb + t) * f] = l;
      if (fabs(l) > lim) bad = 1;
    }
  }
  if (bad) atomicAdd(N.counters + CNT_GROWTH, 1);
}

// C[i,j] -= sum_t A[i,t] * B[j,t] on the lower trapezoid i >= j (global coordinates aligned:
// row i of C and column j of C refer to the same front index origin). 64x64 tiles, 256 threads.
#define TM 64
#define TK 16
__device__ __forceinline__ void tile_syrk(double* __restrict__ C, long long ldc,
                                          const double* __restrict__ A,
                                          const double* __restrict__ Bm, long long ld, int M, int Nn,
                                          int K, int ti, int tj) {
  __shared__ double As[TK][TM + 1];
  __shared__ double Bs[TK][TM + 1];
  const int i0 = ti * TM, j0 = tj * TM;
  if (i0 + TM - 1 < j0) return;  // tile strictly above the diagonal
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < K; k0 += TK) {
    for (int t = threadIdx.x; t < TM * TK; t += 256) {
      int ii = t % TM, kk = t / TM;
      int gi = i0 + ii, gj = j0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < M && gk < K) ? A[gi + (long long)gk * ld] : 0.0;
      Bs[kk][ii] = (gj < Nn && gk < K) ? Bm[gj + (long long)gk * ld] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      double a[4], b[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { a[q] = As[kk][tx + 16 * q]; b[q] = Bs[kk][ty + 16 * q]; }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[q][p] = fma(a[q], b[p], acc[q][p]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int gi = i0 + tx + 16 * q, gj = j0 + ty + 16 * p;
      if (gi < M && gj < Nn && gi >= gj) C[gi + (long long)gj * ldc] -= acc[q][p];
    }
}

// trailing update of the remaining pivot columns after panel jb
__global__ void __launch_bounds__(256) k_big_update(DevSym S, DevNum N, const int* __restrict__ front_list, int jb) {
  const int s = front_list[blockIdx.z];
  const int k = S.sn_start[s + 1] - S.sn_start[s];
  if (jb + NB >= k) return;
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int o = jb + NB;
  const int M = f - o, Nn = k - o;
  if ((int)blockIdx.x * TM >= M || (int)blockIdx.y * TM >= Nn) return;
  double* P = N.L + S.L_off[s];
  const double* Wp = N.W + S.L_off[s];
  tile_syrk(P + o + (long long)o * f, f, P + o + (long long)jb * f, Wp + o + (long long)jb * f, f, M, Nn,
            NB, blockIdx.x, blockIdx.y);
}

// Schur complement: CB -= L21 * (L21 D)^T  (the dense contraction of the front)
__global__ void __launch_bounds__(256) k_big_schur(DevSym S, DevNum N, const int* __restrict__ front_list) {
  const int s = front_list[blockIdx.z];
  const int k = S.sn_start[s + 1] - S.sn_start[s];
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int f = k + r;
  if ((int)blockIdx.x * TM >= r || (int)blockIdx.y * TM >= r) return;
  double* C = N.CB + S.cb_off[s];
  const double* P = N.L + S.L_off[s];
  const double* Wp = N.W + S.L_off[s];
  tile_syrk(C, r, P + k, Wp + k, f, r, r, k, blockIdx.x, blockIdx.y);
}

}  // namespace b200
