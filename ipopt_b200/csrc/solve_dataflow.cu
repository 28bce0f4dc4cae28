// Persistent dataflow triangular solve (forward L, D^-1, backward L^T) -- ONE kernel per sweep.
//
// The level-per-launch version (solve_kernels.cu) spends its time in launch gaps and in single-CTA big
// fronts.  Here every front (or 64-row / 64-column block of a big front) is a TASK in a topologically
// sorted list; persistent CTAs take tasks with an atomic ticket and wait for their producers through
// acquire/release flags in global memory.  A task only ever waits for tasks that precede it in the list,
// and every ticket is held by a resident CTA, so the scheme cannot deadlock.  L is streamed exactly once
// per sweep (HBM-bound, SURVEY.md 8d: 2*8*nnz(L) bytes per right-hand side); no atomics on the data path
// (children -> parent through per-front update vectors gathered by the parent), so results are
// bit-reproducible.  Replaces the vendor back-solve of the reference (MUMPS job=3,
// reference src/Algorithm/LinearSolvers/IpMumpsSolverInterface.cpp:543-583).
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.cuh"

namespace b200 {

#define DF_THREADS 256
#define DF_BLK 64          // block-row / block-column size of the big-front tasks
#define DF_SMALL_SMEM (33 * 32 + 64)   // doubles per warp for a small front

enum { ST_SMALL = 0, ST_MID = 1, ST_BIG_GATHER = 2, ST_BIG_BLOCK = 3 };

struct SolveTask { int type, a, b, c; };

struct DevSolve {
  const SolveTask* tasks;      // forward list
  const SolveTask* tasks_bwd;  // backward list
  int ntasks_fwd, ntasks_bwd;
  const int* bundle;           // front ids of the small bundles
  int* done_f;                 // nsn : epoch when the forward work of a front is complete
  int* done_b;                 // nsn : same for backward
  int* gflag;                  // nsn : big-front gather done
  int* bflag_f;                // per (big front, block): y block published
  int* bflag_b;                // per (big front, block): x block published
  int* bcnt;                   // nsn : finished forward block counter (monotonic)
  int* bcnt_b;                 // nsn : finished backward block counter (monotonic)
  const int* boff;             // nsn : offset of a big front's blocks in bflag_*
  const long long* bigv_off;   // nsn : offset into bigv (f doubles) / bigy (k doubles at the same offset)
  double* bigv;                // assembled+permuted rhs of big fronts
  double* bigy;                // y / x blocks of big fronts in pivoted order
  unsigned long long* ticket;  // [0] fwd, [1] bwd (monotonic)
  const double* binv;          // per (big front, 64-block): Ext[64x64] (rows x pivot cols: [L_dd^-1 ; -L_cd L_dd^-1]) then L_dd^-T-friendly transpose
  const long long* binv_off;   // nsn : offset of a big front's first block in binv (8192 doubles per block), -1 = none
  unsigned long long* tlog;    // optional (debug): 2 timestamps per task, fwd then bwd; nullptr = off
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_eq(const int* p, int epoch) {
  while (ld_acquire(p) != epoch) __nanosleep(32);
}

// ------------------------------------------------------------------------------------------------
// small fronts (order <= 32): one warp per front, panel staged in shared memory
// ------------------------------------------------------------------------------------------------
__device__ void small_fwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                          double* __restrict__ x, double* __restrict__ cbv) {
  const int lane = threadIdx.x & 31;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  double* Ls = sm;            // f x k, ld 33
  double* w = sm + 33 * 32;   // 32
  const int ch0 = S.child_ptr[s], ch1 = S.child_ptr[s + 1];
  // stage the panel first: it does not depend on the children
  const double* __restrict__ P = N.L + S.L_off[s];
  for (int tb = 0; tb < k; tb += 8) {   // 8 independent loads in flight per lane
    double tmp[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) tmp[q] = (lane < f && tb + q < k) ? P[lane + (size_t)(tb + q) * f] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (tb + q < k) Ls[lane + (tb + q) * 33] = tmp[q];
  }
  w[lane] = (lane < k) ? x[c0 + lane] : 0.0;
  for (int q = ch0 + lane; q < ch1; q += 32) wait_eq(V.done_f + S.child_idx[q], epoch);
  __syncwarp();
  for (int q = ch0; q < ch1; ++q) {
    const int c = S.child_idx[q];
    const long long o = S.rows_ptr[c];
    const int rc = (int)(S.rows_ptr[c + 1] - o);
    if (lane < rc) w[S.rel[o + lane]] += __ldcg(cbv + o + lane);
    __syncwarp();
  }
  double v = 0.0;
  if (lane < f) v = (lane < k) ? w[N.lperm[c0 + lane]] : w[lane];
  for (int t = 0; t < k; ++t) {
    const double yt = __shfl_sync(0xffffffffu, v, t);
    if (lane > t && lane < f) v = fma(-Ls[lane + t * 33], yt, v);
  }
  // D^-1 (2x2 partners are neighbouring lanes)
  const double vn = __shfl_down_sync(0xffffffffu, v, 1), vp = __shfl_up_sync(0xffffffffu, v, 1);
  if (lane < k) {
    const int ty = N.ptype[c0 + lane];
    double y;
    if (ty == 1) y = v * N.dinv[c0 + lane];
    else if (ty == 2) y = v * N.dinv[c0 + lane] + vn * N.doff[c0 + lane];
    else y = vp * N.doff[c0 + lane - 1] + v * N.dinv[c0 + lane];
    x[c0 + lane] = y;
  } else if (lane < f) cbv[ro + lane - k] = v;
  __syncwarp();
  if (lane == 0) { st_release(V.done_f + s, epoch); }
}

__device__ void small_bwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                          double* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  double* Ls = sm;
  const double* __restrict__ P = N.L + S.L_off[s];
  for (int tb = 0; tb < k; tb += 8) {   // 8 independent loads in flight per lane
    double tmp[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) tmp[q] = (lane < f && tb + q < k) ? P[lane + (size_t)(tb + q) * f] : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (tb + q < k) Ls[lane + (tb + q) * 33] = tmp[q];
  }
  const int par = S.sn_parent[s];
  if (par >= 0 && lane == 0) wait_eq(V.done_b + par, epoch);
  __syncwarp();
  double v = 0.0;  // lane i holds entry i of [D^-1 y ; x(rows)]
  if (lane < k) v = x[c0 + lane];
  else if (lane < f) v = __ldcg(x + S.rows[ro + lane - k]);
  // columns from the last to the first: v_t -= sum_{i>t} L[i,t] v_i.  Lane t owns column t.
  for (int i = f - 1; i >= 1; --i) {
    const double vi = __shfl_sync(0xffffffffu, v, i);
    if (lane < i && lane < k) v = fma(-Ls[i + lane * 33], vi, v);
  }
  if (lane < k) x[c0 + N.lperm[c0 + lane]] = v;
  __syncwarp();
  if (lane == 0) { st_release(V.done_b + s, epoch); }
}

// ------------------------------------------------------------------------------------------------
// fronts of order 33..64: one warp per front, two rows per lane (lane, lane+32), no shared-memory staging:
// the L loads do not depend on the running vector, so they are issued ahead of the shuffle chain.
// smem per warp: w[64]
// ------------------------------------------------------------------------------------------------
__device__ void w64_fwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                        double* __restrict__ x, double* __restrict__ cbv) {
  const int lane = threadIdx.x & 31;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  double* w = sm;  // 64
  const int ch0 = S.child_ptr[s], ch1 = S.child_ptr[s + 1];
  const int i0 = lane, i1 = lane + 32;
  w[i0] = (i0 < k) ? x[c0 + i0] : 0.0;
  w[i1] = (i1 < k) ? x[c0 + i1] : 0.0;
  for (int q = ch0 + lane; q < ch1; q += 32) wait_eq(V.done_f + S.child_idx[q], epoch);
  __syncwarp();
  for (int q = ch0; q < ch1; ++q) {
    const int c = S.child_idx[q];
    const long long o = S.rows_ptr[c];
    const int rc = (int)(S.rows_ptr[c + 1] - o);
    for (int t = lane; t < rc; t += 32) w[S.rel[o + t]] += __ldcg(cbv + o + t);
    __syncwarp();
  }
  double v0 = 0.0, v1 = 0.0;
  if (i0 < f) v0 = (i0 < k) ? w[N.lperm[c0 + i0]] : w[i0];
  if (i1 < f) v1 = (i1 < k) ? w[N.lperm[c0 + i1]] : w[i1];
  const double* __restrict__ P = N.L + S.L_off[s];
  for (int tb = 0; tb < k; tb += 8) {
    double l0[8], l1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tb + q;
      l0[q] = (t < k && i0 > t && i0 < f) ? P[i0 + (size_t)t * f] : 0.0;
      l1[q] = (t < k && i1 > t && i1 < f) ? P[i1 + (size_t)t * f] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tb + q;
      if (t < k) {
        const double yt = (t < 32) ? __shfl_sync(0xffffffffu, v0, t) : __shfl_sync(0xffffffffu, v1, t - 32);
        v0 = fma(-l0[q], yt, v0);
        v1 = fma(-l1[q], yt, v1);
      }
    }
  }
  // D^-1: partner values through shared memory (w is free now)
  __syncwarp();
  w[i0] = v0; w[i1] = v1;
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int i = lane + 32 * h;
    const double v = h ? v1 : v0;
    if (i < k) {
      const int ty = N.ptype[c0 + i];
      double y;
      if (ty == 1) y = v * N.dinv[c0 + i];
      else if (ty == 2) y = v * N.dinv[c0 + i] + w[i + 1] * N.doff[c0 + i];
      else y = w[i - 1] * N.doff[c0 + i - 1] + v * N.dinv[c0 + i];
      x[c0 + i] = y;
    } else if (i < f) cbv[ro + i - k] = v;
  }
  __syncwarp();
  if (lane == 0) { st_release(V.done_f + s, epoch); }
}

__device__ void w64_bwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                        double* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  const int i0 = lane, i1 = lane + 32;
  const double* __restrict__ P = N.L + S.L_off[s];
  const int par = S.sn_parent[s];
  if (par >= 0 && lane == 0) wait_eq(V.done_b + par, epoch);
  __syncwarp();
  double v0 = 0.0, v1 = 0.0;  // entries i0 / i1 of [D^-1 y ; x(rows)]
  if (i0 < k) v0 = x[c0 + i0]; else if (i0 < f) v0 = __ldcg(x + S.rows[ro + i0 - k]);
  if (i1 < k) v1 = x[c0 + i1]; else if (i1 < f) v1 = __ldcg(x + S.rows[ro + i1 - k]);
  // columns from the last to the first: v_t -= sum_{i>t} L[i,t] v_i  (column read coalesced, warp-sum)
  for (int tb = k - 1; tb >= 0; tb -= 8) {
    double l0[8], l1[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tb - q;
      l0[q] = (t >= 0 && i0 > t && i0 < f) ? P[i0 + (size_t)t * f] : 0.0;
      l1[q] = (t >= 0 && i1 > t && i1 < f) ? P[i1 + (size_t)t * f] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tb - q;
      if (t >= 0) {
        double part = fma(l0[q], v0, l1[q] * v1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (t < 32) { if (lane == t) v0 -= part; } else { if (lane == t - 32) v1 -= part; }
      }
    }
  }
  if (i0 < k) x[c0 + N.lperm[c0 + i0]] = v0;
  if (i1 < k) x[c0 + N.lperm[c0 + i1]] = v1;
  __syncwarp();
  if (lane == 0) { st_release(V.done_b + s, epoch); }
}

// ------------------------------------------------------------------------------------------------
// mid fronts (33 .. mid_max): one CTA, blocked by 32 (same algorithm as solve_kernels.cu)
// smem: v[f] | w[f] | Lb[32*33]
// ------------------------------------------------------------------------------------------------
__device__ void mid_fwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                        double* __restrict__ x, double* __restrict__ cbv) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]), f = k + r;
  double* v = sm;
  double* w = sm + f;
  double* Lb = sm + 2 * f;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int ch0 = S.child_ptr[s], ch1 = S.child_ptr[s + 1];
  for (int q = ch0 + tid; q < ch1; q += nt) wait_eq(V.done_f + S.child_idx[q], epoch);
  for (int i = tid; i < f; i += nt) w[i] = (i < k) ? x[c0 + i] : 0.0;
  __syncthreads();
  for (int q = ch0; q < ch1; ++q) {
    const int c = S.child_idx[q];
    const long long o = S.rows_ptr[c];
    const int rc = (int)(S.rows_ptr[c + 1] - o);
    for (int t = tid; t < rc; t += nt) w[S.rel[o + t]] += __ldcg(cbv + o + t);
    __syncthreads();
  }
  const int* __restrict__ lp = N.lperm + c0;
  for (int i = tid; i < f; i += nt) v[i] = (i < k) ? w[lp[i]] : w[i];
  __syncthreads();
  const double* __restrict__ P = N.L + S.L_off[s];
  for (int t0 = 0; t0 < k; t0 += 32) {
    const int nb = min(32, k - t0);
    for (int t = tid; t < nb * nb; t += nt) {
      int i = t % nb, q = t / nb;
      Lb[i + q * 33] = P[(t0 + i) + (size_t)(t0 + q) * f];
    }
    __syncthreads();
    if (warp == 0) {
      double yi = (lane < nb) ? v[t0 + lane] : 0.0;
      for (int q = 0; q < nb; ++q) {
        double yq = __shfl_sync(0xffffffffu, yi, q);
        if (lane > q && lane < nb) yi = fma(-Lb[lane + q * 33], yq, yi);
      }
      if (lane < nb) v[t0 + lane] = yi;
    }
    __syncthreads();
    for (int i = t0 + nb + tid; i < f; i += nt) {
      double acc = 0.0;
#pragma unroll 8
      for (int q = 0; q < nb; ++q) acc = fma(P[i + (size_t)(t0 + q) * f], v[t0 + q], acc);
      v[i] -= acc;
    }
    __syncthreads();
  }
  for (int t = tid; t < k; t += nt) {
    const int ty = N.ptype[c0 + t];
    double y;
    if (ty == 1) y = v[t] * N.dinv[c0 + t];
    else if (ty == 2) y = v[t] * N.dinv[c0 + t] + v[t + 1] * N.doff[c0 + t];
    else y = v[t - 1] * N.doff[c0 + t - 1] + v[t] * N.dinv[c0 + t];
    x[c0 + t] = y;
  }
  double* __restrict__ out = cbv + S.rows_ptr[s];
  for (int i = tid; i < r; i += nt) out[i] = v[k + i];
  __syncthreads();
  if (tid == 0) { st_release(V.done_f + s, epoch); }
}

__device__ void mid_bwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                        double* __restrict__ x) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  double* v = sm;
  double* Lb = sm + 2 * f;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  const int par = S.sn_parent[s];
  if (par >= 0 && tid == 0) wait_eq(V.done_b + par, epoch);
  __syncthreads();
  for (int i = tid; i < f; i += nt) v[i] = (i < k) ? x[c0 + i] : __ldcg(x + S.rows[ro + (i - k)]);
  __syncthreads();
  const double* __restrict__ P = N.L + S.L_off[s];
  const int nblk = (k + 31) / 32;
  for (int b = nblk - 1; b >= 0; --b) {
    const int t0 = b * 32, nb = min(32, k - t0);
    for (int t = tid; t < nb * nb; t += nt) {
      int i = t % nb, q = t / nb;
      Lb[i + q * 33] = P[(t0 + i) + (size_t)(t0 + q) * f];
    }
    for (int q = warp; q < nb; q += nwarp) {
      const double* col = P + (size_t)(t0 + q) * f;
      double acc = 0.0;
      for (int i = t0 + nb + lane; i < f; i += 32) acc = fma(col[i], v[i], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) v[t0 + q] -= acc;
    }
    __syncthreads();
    if (warp == 0) {
      double zi = (lane < nb) ? v[t0 + lane] : 0.0;
      for (int q = nb - 1; q >= 0; --q) {
        double zq = __shfl_sync(0xffffffffu, zi, q);
        if (lane < q) zi = fma(-Lb[q + lane * 33], zq, zi);
      }
      if (lane < nb) v[t0 + lane] = zi;
    }
    __syncthreads();
  }
  const int* __restrict__ lp = N.lperm + c0;
  for (int t = tid; t < k; t += nt) x[c0 + lp[t]] = v[t];
  __syncthreads();
  if (tid == 0) { st_release(V.done_b + s, epoch); }
}

// ------------------------------------------------------------------------------------------------
// big fronts: gather task + one task per 64-row block (forward) / 64-column block (backward)
// ------------------------------------------------------------------------------------------------
__device__ void big_gather(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch,
                           const double* __restrict__ x, const double* __restrict__ cbv) {
  // w = [x(cols) ; 0] + scatter(children update vectors), then the in-front pivot permutation; result in bigv.
  // Children write disjoint... no: two children may hit the same parent row, so children are applied one
  // after the other (deterministic), each child fully parallel.
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]), f = k + r;
  const int tid = threadIdx.x, nt = blockDim.x;
  double* w = V.bigv + V.bigv_off[s];      // scratch (pre-permutation) lives in bigy, result in bigv
  double* tmp = V.bigy + V.bigv_off[s];    // f doubles available (bigy has the same per-front extent)
  const int ch0 = S.child_ptr[s], ch1 = S.child_ptr[s + 1];
  for (int q = ch0 + tid; q < ch1; q += nt) wait_eq(V.done_f + S.child_idx[q], epoch);
  for (int i = tid; i < f; i += nt) tmp[i] = (i < k) ? x[c0 + i] : 0.0;
  __syncthreads();
  for (int q = ch0; q < ch1; ++q) {
    const int c = S.child_idx[q];
    const long long o = S.rows_ptr[c];
    const int rc = (int)(S.rows_ptr[c + 1] - o);
    for (int t = tid; t < rc; t += nt) tmp[S.rel[o + t]] += __ldcg(cbv + o + t);
    __syncthreads();
  }
  const int* __restrict__ lp = N.lperm + c0;
  for (int i = tid; i < f; i += nt) w[i] = (i < k) ? tmp[lp[i]] : tmp[i];
  __syncthreads();
  if (tid == 0) { st_release(V.gflag + s, epoch); }
}

// forward block row b of big front s: rows [64b, min(f, 64b+64))
// smem: ys[64] | part[4*64] | Lsq[64*65]
__device__ void big_fwd_block(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int b, int epoch,
                              double* sm, double* __restrict__ x, double* __restrict__ cbv) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]), f = k + r;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31, warp = tid >> 5;
  double* ys = sm;
  double* part = sm + 64;
  double* Lsq = sm + 64 + 256;
  const int r0 = b * DF_BLK, nrow = min(DF_BLK, f - r0);
  const int ncolblk_left = min(b, (k + DF_BLK - 1) / DF_BLK);   // column blocks strictly left of the diagonal block
  const double* __restrict__ P = N.L + S.L_off[s];
  const int* bf = V.bflag_f + V.boff[s];
  const double* yb = V.bigy + V.bigv_off[s];
  double acc = 0.0;
  const int row = r0 + tx;
  // everything that does not depend on the y blocks is fetched first: the diagonal block and (after the gather
  // flag, which is long set by the time the left blocks arrive) this block's slice of the assembled rhs
  const int ndiag = max(0, min(k - r0, DF_BLK));
  if (ndiag > 0) {
    for (int t = tid; t < nrow * ndiag; t += blockDim.x) {
      int i = t % nrow, q = t / nrow;
      Lsq[i + q * 65] = P[(r0 + i) + (size_t)(r0 + q) * f];
    }
  }
  if (tid == 0) wait_eq(V.gflag + s, epoch);
  __syncthreads();
  double wmine = 0.0;
  if (tid < nrow) wmine = __ldcg(V.bigv + V.bigv_off[s] + r0 + tid);
  // precomputed inverse of the diagonal block (k_big_blockinv): the in-block substitution becomes a 64x64 GEMV
  const bool use_inv = (V.binv != nullptr) && V.binv_off[s] >= 0 && ndiag > 0;
  double ext[16];
  if (use_inv) {
    const double* __restrict__ E = V.binv + V.binv_off[s] + (size_t)b * 8192;
#pragma unroll
    for (int q = 0; q < 16; ++q) ext[q] = E[tx + (ty + 4 * q) * 64];
  }
  // software pipeline: the L tile of column block c+1 is in flight while block c is consumed
  double ltn[16];
  if (ncolblk_left > 0) {
    const int nc0 = min(DF_BLK, k);
    const double* base = P + row;
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nrow && t < nc0) ? base[(size_t)t * f] : 0.0; }
  }
  for (int c = 0; c < ncolblk_left; ++c) {
    const int t0 = c * DF_BLK, ncol = min(DF_BLK, k - t0);
    double lt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
    if (c + 1 < ncolblk_left) {
      const int t0n = t0 + DF_BLK, ncn = min(DF_BLK, k - t0n);
      const double* base = P + row + (size_t)t0n * f;
#pragma unroll
      for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nrow && t < ncn) ? base[(size_t)t * f] : 0.0; }
    }
    if (tid == 0) wait_eq(bf + c, epoch);
    __syncthreads();
    if (tid < ncol) ys[tid] = __ldcg(yb + t0 + tid);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; if (t < ncol) acc = fma(lt[q], ys[t], acc); }
    __syncthreads();
  }
  part[ty * 64 + tx] = acc;
  __syncthreads();
  if (tid < 64) ys[tid] = (tid < nrow) ? wmine - (part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid]) : 0.0;
  __syncthreads();
  if (use_inv) {
    double p2 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; if (t < ndiag) p2 = fma(ext[q], ys[t], p2); }
    __syncthreads();                       // everyone has read ys
    part[ty * 64 + tx] = p2;
    __syncthreads();
    if (tid < 64) {
      const double g = part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid];
      const double mine = ys[tid];
      __syncwarp();
      ys[tid] = (tid < ndiag) ? g : mine + g;   // pivot rows: y = L_dd^-1 v ; contribution rows of the block: v - L_cd y
    }
  } else if (warp == 0) {
    double v0 = ys[lane], v1 = ys[lane + 32];
#pragma unroll 8
    for (int t = 0; t < ndiag; ++t) {
      const double yt = (t < 32) ? __shfl_sync(0xffffffffu, v0, t) : __shfl_sync(0xffffffffu, v1, t - 32);
      if (lane > t) v0 = fma(-Lsq[lane + t * 65], yt, v0);
      if (lane + 32 > t) v1 = fma(-Lsq[lane + 32 + t * 65], yt, v1);
    }
    ys[lane] = v0; ys[lane + 32] = v1;
  }
  __syncthreads();
  if (tid < nrow) {
    const int i = r0 + tid;
    if (i < k) {
      V.bigy[V.bigv_off[s] + i] = ys[tid];   // y block for the rows below (pivoted order)
      const int ty2 = N.ptype[c0 + i];
      double y;
      if (ty2 == 1) y = ys[tid] * N.dinv[c0 + i];
      else if (ty2 == 2) y = ys[tid] * N.dinv[c0 + i] + ys[tid + 1] * N.doff[c0 + i];
      else y = ys[tid - 1] * N.doff[c0 + i - 1] + ys[tid] * N.dinv[c0 + i];
      x[c0 + i] = y;
    } else cbv[S.rows_ptr[s] + (i - k)] = ys[tid];
  }
  __syncthreads();
  if (tid == 0) {
    if (r0 < k) st_release((int*)bf + b, epoch);
    const int nblk = (f + DF_BLK - 1) / DF_BLK;
    const int old = atomicAdd(V.bcnt + s, 1);
    if ((old + 1) % nblk == 0) { st_release(V.done_f + s, epoch); }
  }
}

// backward block column b of big front s: columns [64b, min(k, 64b+64))
// smem: xs[64] | red[DF_THREADS*16 -> done with shuffles] ...
__device__ void big_bwd_block(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int b, int epoch,
                              double* sm, double* __restrict__ x) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6, lane = tid & 31, warp = tid >> 5;
  double* xs = sm;                 // 64
  double* colacc = sm + 64;        // 64
  double* red = sm + 128;          // 8 warps x 16
  double* Lsq = sm + 128 + 128;    // 64 x 65
  const int t0 = b * DF_BLK, ncol = min(DF_BLK, k - t0), t1 = t0 + ncol;
  const int nkb = (k + DF_BLK - 1) / DF_BLK;
  const double* __restrict__ P = N.L + S.L_off[s];
  const int* bf = V.bflag_b + V.boff[s];
  double pacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) pacc[q] = 0.0;
  // stage the diagonal block first (independent of everything we wait for)
  for (int t = tid; t < ncol * ncol; t += blockDim.x) {
    int i = t % ncol, q = t / ncol;
    Lsq[i + q * 65] = P[(t0 + i) + (size_t)(t0 + q) * f];
  }
  const int par = S.sn_parent[s];
  if (par >= 0 && tid == 0) wait_eq(V.done_b + par, epoch);
  __syncthreads();
  // rows below the block are processed in chunks of 64: first the contribution-block rows (ancestors'
  // final values), then this front's later pivot blocks from the last one down as they get published.
  const int nchunk_cb = (r + DF_BLK - 1) / DF_BLK;
  const int nchunk = nchunk_cb + (nkb - 1 - b);
  // chunk geometry: first the contribution-block rows, then the later pivot blocks from the last one down
  auto chunk_geom = [&](int ch, int& rbase, int& nr, int& c) {
    if (ch < nchunk_cb) { c = -1; rbase = k + ch * DF_BLK; nr = min(DF_BLK, f - rbase); }
    else { c = nkb - 1 - (ch - nchunk_cb); rbase = c * DF_BLK; nr = min(DF_BLK, k - rbase); }
  };
  double ltn[16];
  if (nchunk > 0) {
    int rbase, nr, c;
    chunk_geom(0, rbase, nr, c);
    const double* base = P + (rbase + tx) + (size_t)t0 * f;
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr && t < ncol) ? base[(size_t)t * f] : 0.0; }
  }
  for (int ch = 0; ch < nchunk; ++ch) {
    int rbase, nr, c;
    chunk_geom(ch, rbase, nr, c);
    double lt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
    if (ch + 1 < nchunk) {   // next tile in flight while this chunk is consumed (it does not depend on x)
      int rb2, nr2, c2;
      chunk_geom(ch + 1, rb2, nr2, c2);
      const double* base = P + (rb2 + tx) + (size_t)t0 * f;
#pragma unroll
      for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr2 && t < ncol) ? base[(size_t)t * f] : 0.0; }
    }
    if (c < 0) {
      if (tid < nr) xs[tid] = __ldcg(x + S.rows[ro + (rbase - k) + tid]);
    } else {
      if (tid == 0) wait_eq(bf + c, epoch);
      __syncthreads();
      if (tid < nr) xs[tid] = __ldcg(V.bigy + V.bigv_off[s] + rbase + tid);
    }
    __syncthreads();
    if (tx < nr) {
      const double xi = xs[tx];
#pragma unroll
      for (int q = 0; q < 16; ++q) pacc[q] = fma(lt[q], xi, pacc[q]);
    }
    __syncthreads();
  }
  // reduce pacc over the 64 row-threads of each ty group: warp shuffle then across the two half-warps
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    double a = pacc[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) red[warp * 16 + q] = a;
  }
  __syncthreads();
  if (tid < 64) {
    // column t = tid: ty = t & 3, q = t >> 2 ; warps 2*ty and 2*ty+1 hold the two halves
    const int tyc = tid & 3, q = tid >> 2;
    colacc[tid] = red[(2 * tyc) * 16 + q] + red[(2 * tyc + 1) * 16 + q];
  }
  __syncthreads();
  const bool use_inv = (V.binv != nullptr) && V.binv_off[s] >= 0;
  if (use_inv) {
    // z_t = sum_{i >= t} Linv[i,t] * rhs_i  with the transposed copy (coalesced over t)
    const double* __restrict__ ET = V.binv + V.binv_off[s] + (size_t)b * 8192 + 4096;
    if (tid < 64) xs[tid] = (tid < ncol) ? x[c0 + t0 + tid] - colacc[tid] : 0.0;
    __syncthreads();
    double p2 = 0.0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int i = ty + 4 * q; if (i < ncol) p2 = fma(ET[tx + i * 64], xs[i], p2); }
    __syncthreads();
    red[0] = 0.0;   // (red is free again)
    double* part2 = Lsq;   // 256 doubles of scratch
    part2[ty * 64 + tx] = p2;
    __syncthreads();
    if (tid < 64) xs[tid] = part2[tid] + part2[64 + tid] + part2[128 + tid] + part2[192 + tid];
  } else if (warp == 0) {
    double z0 = (lane < ncol) ? x[c0 + t0 + lane] - colacc[lane] : 0.0;
    double z1 = (lane + 32 < ncol) ? x[c0 + t0 + lane + 32] - colacc[lane + 32] : 0.0;
#pragma unroll 8
    for (int q = ncol - 1; q >= 0; --q) {
      const double zq = (q < 32) ? __shfl_sync(0xffffffffu, z0, q) : __shfl_sync(0xffffffffu, z1, q - 32);
      if (lane < q) z0 = fma(-Lsq[q + lane * 65], zq, z0);
      if (lane + 32 < q) z1 = fma(-Lsq[q + (lane + 32) * 65], zq, z1);
    }
    xs[lane] = z0; xs[lane + 32] = z1;
  }
  __syncthreads();
  if (tid < ncol) {
    V.bigy[V.bigv_off[s] + t0 + tid] = xs[tid];                 // pivoted order, for the blocks to the left
    x[c0 + N.lperm[c0 + t0 + tid]] = xs[tid];                   // final value (permutation is panel-local)
  }
  __syncthreads();
  if (tid == 0) {
    st_release((int*)bf + b, epoch);
    const int old = atomicAdd(V.bcnt_b + s, 1);
    if ((old + 1) % nkb == 0) { st_release(V.done_b + s, epoch); }
  }
  (void)t1;
}

// ------------------------------------------------------------------------------------------------
// Inverses of the 64x64 diagonal blocks of the big fronts' L11 (run once per factorisation, all blocks in parallel):
//   Ext (64 x 64, ld 64): rows = rows of block b of the front, cols = its pivot columns:
//        pivot rows      : L_dd^-1               (unit lower triangular)
//        rows below k    : -L_cd * L_dd^-1       (contribution rows that share the block with the last pivots)
//   ET  (64 x 64, ld 64): ET[t + i*64] = L_dd^-1[i][t]   (transposed copy for the backward sweep)
// One CTA of 64 threads per (front, block); list = pairs (front, block).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_big_blockinv(DevSym S, DevNum N, const int* __restrict__ pairs,
                                                     const long long* __restrict__ binv_off, double* __restrict__ binv) {
  extern __shared__ double binv_sm[];
  double* Ls = binv_sm;
  double* Xs = binv_sm + 64 * 65;
  const int s = pairs[2 * blockIdx.x], b = pairs[2 * blockIdx.x + 1];
  const int k = S.sn_start[s + 1] - S.sn_start[s];
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int r0 = b * 64, nrow = min(64, f - r0), nd = min(64, k - r0);
  const double* __restrict__ P = N.L + S.L_off[s];
  double* __restrict__ E = binv + binv_off[s] + (size_t)b * 8192;
  const int tid = threadIdx.x;
  for (int t = tid; t < 64 * 64; t += 64) {
    const int i = t & 63, q = t >> 6;
    Ls[i + q * 65] = (i < nrow && q < nd) ? P[(r0 + i) + (size_t)(r0 + q) * f] : 0.0;
    Xs[i + q * 65] = 0.0;
  }
  __syncthreads();
  // thread j computes column j of X = L_dd^-1 by forward substitution (unit diagonal)
  if (tid < nd) {
    const int j = tid;
    Xs[j + j * 65] = 1.0;
    for (int i = j + 1; i < nd; ++i) {
      double acc = 0.0;
      for (int t = j; t < i; ++t) acc = fma(Ls[i + t * 65], Xs[t + j * 65], acc);
      Xs[i + j * 65] = -acc;
    }
  }
  __syncthreads();
  // rows of the block below the pivots: -L_cd * X
  if (tid >= nd && tid < nrow) {
    const int i = tid;
    for (int j = 0; j < nd; ++j) {
      double acc = 0.0;
      for (int t = j; t < nd; ++t) acc = fma(Ls[i + t * 65], Xs[t + j * 65], acc);
      Xs[i + j * 65] = -acc;
    }
  }
  __syncthreads();
  for (int t = tid; t < 64 * 64; t += 64) {
    const int i = t & 63, q = t >> 6;
    E[i + q * 64] = Xs[i + q * 65];
    E[4096 + i + q * 64] = (i < nd && q < nd) ? Xs[q + i * 65] : 0.0;   // ET[t=i + row q*64] = X[q][i]
  }
}

// ------------------------------------------------------------------------------------------------
template <bool FWD>
__global__ void __launch_bounds__(DF_THREADS) k_solve_dataflow(DevSym S, DevNum N, DevSolve V, int epoch,
                                                                unsigned long long ticket_base,
                                                                double* __restrict__ x, double* __restrict__ cbv) {
  extern __shared__ double sm[];
  __shared__ unsigned long long s_ticket;
  const SolveTask* tasks = FWD ? V.tasks : V.tasks_bwd;
  const int ntasks = FWD ? V.ntasks_fwd : V.ntasks_bwd;
  while (true) {
    if (threadIdx.x == 0) s_ticket = atomicAdd(V.ticket + (FWD ? 0 : 1), 1ull) - ticket_base;
    __syncthreads();
    const unsigned long long tk = s_ticket;
    __syncthreads();
    if (tk >= (unsigned long long)ntasks) return;
    const SolveTask T = tasks[tk];
    unsigned long long t_start = 0;
    if (V.tlog && threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));
    if (T.type == ST_SMALL) {
      const int w = threadIdx.x >> 5;
      if (w < T.b) {
        const int s = V.bundle[T.a + w];
        double* wsm = sm + (size_t)w * DF_SMALL_SMEM;
        const int fs = (S.sn_start[s + 1] - S.sn_start[s]) + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
        if (fs <= 32) { if (FWD) small_fwd(S, N, V, s, epoch, wsm, x, cbv); else small_bwd(S, N, V, s, epoch, wsm, x); }
        else { if (FWD) w64_fwd(S, N, V, s, epoch, wsm, x, cbv); else w64_bwd(S, N, V, s, epoch, wsm, x); }
      }
    } else if (T.type == ST_MID) {
      if (FWD) mid_fwd(S, N, V, T.a, epoch, sm, x, cbv); else mid_bwd(S, N, V, T.a, epoch, sm, x);
    } else if (T.type == ST_BIG_GATHER) {
      big_gather(S, N, V, T.a, epoch, x, cbv);
    } else {
      if (FWD) big_fwd_block(S, N, V, T.a, T.b, epoch, sm, x, cbv); else big_bwd_block(S, N, V, T.a, T.b, epoch, sm, x);
    }
    __syncthreads();
    if (V.tlog && threadIdx.x == 0) {
      unsigned long long t_end;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
      unsigned long long* rec = V.tlog + 2 * ((FWD ? 0 : (unsigned long long)V.ntasks_fwd) + tk);
      rec[0] = t_start; rec[1] = t_end;
    }
  }
}

}  // namespace b200
