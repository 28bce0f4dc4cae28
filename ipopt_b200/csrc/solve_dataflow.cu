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
  const double* linv;          // explicit inverses of the big fronts' pivot blocks L11 (K64 x K64 each, see k_linv_*)
  const long long* linv_off;   // nsn : offset of a big front's inverse in linv, -1 = none
  unsigned long long* tlog;    // optional (debug): 2 timestamps per task, fwd then bwd; nullptr = off
  int opts;                    // bit 0: gather through global memory (debug / comparison)
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
__device__ void big_gather_global(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch,
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

#define DF_STAGE 8192      // doubles of the shared staging area (vectors are staged in chunks of this many rows)
#define DF_GATHER_MAXCH 256

// Same result as big_gather_global with the assembled vector held in shared memory and the children's update
// vectors software-pipelined (the (rel, value) pair of the next chunk is in flight while the current one is added;
// a barrier only between children, whose targets may overlap).  Falls back to the global version for fronts that do
// not fit (order > DF_STAGE or more than DF_GATHER_MAXCH children).
// smem: tmp[DF_STAGE] | meta_o[DF_GATHER_MAXCH] (long long) | meta_rc[DF_GATHER_MAXCH] (int)
__device__ void big_gather(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                           const double* __restrict__ x, const double* __restrict__ cbv) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]), f = k + r;
  const int ch0 = S.child_ptr[s], nch = S.child_ptr[s + 1] - ch0;
  if (f > DF_STAGE || nch > DF_GATHER_MAXCH || (V.opts & 1)) { big_gather_global(S, N, V, s, epoch, x, cbv); return; }
  const int tid = threadIdx.x, nt = blockDim.x;
  double* tmp = sm;
  long long* meta_o = reinterpret_cast<long long*>(sm + DF_STAGE);
  int* meta_rc = reinterpret_cast<int*>(meta_o + DF_GATHER_MAXCH);
  for (int q = tid; q < nch; q += nt) {
    const int c = S.child_idx[ch0 + q];
    const long long o = S.rows_ptr[c];
    meta_o[q] = o;
    meta_rc[q] = (int)(S.rows_ptr[c + 1] - o);
    wait_eq(V.done_f + c, epoch);
  }
  for (int i = tid; i < f; i += nt) tmp[i] = (i < k) ? x[c0 + i] : 0.0;
  __syncthreads();
  {
    int q = 0, t = tid;
    int idx = 0; double val = 0.0; bool ok = false;
    if (nch > 0) {
      ok = t < meta_rc[0];
      if (ok) { idx = S.rel[meta_o[0] + t]; val = __ldcg(cbv + meta_o[0] + t); }
    }
    while (q < nch) {
      const int cidx = idx; const double cval = val; const bool cok = ok;
      int nq = q, ntt = t + nt;
      if (ntt - tid >= meta_rc[q]) { nq = q + 1; ntt = tid; }   // (uniform: depends on the chunk base only)
      ok = false;
      if (nq < nch) {
        const long long o = meta_o[nq];
        ok = ntt < meta_rc[nq];
        if (ok) { idx = S.rel[o + ntt]; val = __ldcg(cbv + o + ntt); }
      }
      if (cok) tmp[cidx] += cval;
      if (nq != q) __syncthreads();
      q = nq; t = ntt;
    }
  }
  double* w = V.bigv + V.bigv_off[s];
  const int* __restrict__ lp = N.lperm + c0;
  for (int i = tid; i < f; i += nt) w[i] = (i < k) ? tmp[lp[i]] : tmp[i];
  __syncthreads();
  if (tid == 0) { st_release(V.gflag + s, epoch); }
}

// Big fronts use the EXPLICIT inverse of their unit-lower-triangular pivot block L11 (k_linv_* below, computed once
// per factorisation): the in-front recurrences  y = L11^-1 w  and  x = L11^-T t  become block GEMVs whose 64-row /
// 64-column blocks are independent tasks -- no chain of nkb dependent steps per front.  Linv is stored K64 x K64
// (K64 = 64*ceil(k/64), column-major, zero-padded, upper block triangle never read).
__device__ __forceinline__ void wait_flags(const int* flags, int first, int last, int epoch) {
  // flags[first..last) all equal to epoch (one flag per thread, in parallel), then a CTA barrier
  for (int c = first + (int)threadIdx.x; c < last; c += blockDim.x) wait_eq(flags + c, epoch);
  __syncthreads();
}

// forward, pivot block b (rows [64b, 64b+64) of the pivoted front): y_b = sum_{c<=b} Linv[b,c] w_c
// smem: stage[DF_STAGE] | part[256] | ys[64]
__device__ void big_fwd_piv(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int b, int epoch,
                            double* sm, double* __restrict__ x) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int nkb = (k + DF_BLK - 1) / DF_BLK;
  const long long K64 = (long long)nkb * DF_BLK;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* part = sm + DF_STAGE;
  double* ys = part + 256;
  const int t0 = b * DF_BLK, nrow = min(DF_BLK, k - t0);
  const double* __restrict__ Li = V.linv + V.linv_off[s] + (t0 + tx);
  const double* __restrict__ w = V.bigv + V.bigv_off[s];
  // first tile in flight before the gather flag is seen (Linv does not depend on the right-hand side)
  double ltn[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) ltn[q] = Li[(long long)(ty + 4 * q) * K64];
  if (tid == 0) wait_eq(V.gflag + s, epoch);
  __syncthreads();
  double acc = 0.0;
  const int ntile = b + 1;
  for (int base = 0; base < ntile; base += DF_STAGE / DF_BLK) {
    const int lim = min(ntile, base + DF_STAGE / DF_BLK);
    __syncthreads();
    for (int i = base * DF_BLK + tid; i < lim * DF_BLK; i += blockDim.x) stage[i - base * DF_BLK] = (i < k) ? __ldcg(w + i) : 0.0;
    __syncthreads();
    for (int c = base; c < lim; ++c) {
      double lt[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
      if (c + 1 < ntile) {
        const double* nx = Li + (long long)(c + 1) * DF_BLK * K64;
#pragma unroll
        for (int q = 0; q < 16; ++q) ltn[q] = nx[(long long)(ty + 4 * q) * K64];
      }
      const double* wc = stage + (c - base) * DF_BLK + ty;
#pragma unroll
      for (int q = 0; q < 16; ++q) acc = fma(lt[q], wc[4 * q], acc);
    }
  }
  part[ty * 64 + tx] = acc;
  __syncthreads();
  if (tid < 64) ys[tid] = (tid < nrow) ? part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid] : 0.0;
  __syncthreads();
  if (tid < nrow) {
    const int i = t0 + tid;
    V.bigy[V.bigv_off[s] + i] = ys[tid];   // y block for the contribution rows (pivoted order)
    const int ty2 = N.ptype[c0 + i];
    double y;
    if (ty2 == 1) y = ys[tid] * N.dinv[c0 + i];
    else if (ty2 == 2) y = ys[tid] * N.dinv[c0 + i] + ys[tid + 1] * N.doff[c0 + i];
    else y = ys[tid - 1] * N.doff[c0 + i - 1] + ys[tid] * N.dinv[c0 + i];
    x[c0 + i] = y;
  }
  __syncthreads();
  if (tid == 0) {
    st_release(V.bflag_f + V.boff[s] + b, epoch);
    const int nblk = nkb + (f - k + DF_BLK - 1) / DF_BLK;
    const int old = atomicAdd(V.bcnt + s, 1);
    if ((old + 1) % nblk == 0) { st_release(V.done_f + s, epoch); }
  }
}

// forward, contribution rows [k + 64j, ...): update vector = w - L21[rows, :] y
__device__ void big_fwd_cb(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int j, int epoch,
                           double* sm, double* __restrict__ cbv) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int nkb = (k + DF_BLK - 1) / DF_BLK;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* part = sm + DF_STAGE;
  const int rbase = k + j * DF_BLK, nr = min(DF_BLK, f - rbase);
  const double* __restrict__ P = N.L + S.L_off[s] + (rbase + tx);
  const double* __restrict__ yb = V.bigy + V.bigv_off[s];
  double ltn[16];
  {
    const int nc0 = min(DF_BLK, k);
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr && t < nc0) ? P[(size_t)t * f] : 0.0; }
  }
  if (tid == 0) wait_eq(V.gflag + s, epoch);
  __syncthreads();
  const double wmine = (tid < nr) ? __ldcg(V.bigv + V.bigv_off[s] + rbase + tid) : 0.0;
  wait_flags(V.bflag_f + V.boff[s], 0, nkb, epoch);
  double acc = 0.0;
  for (int base = 0; base < nkb; base += DF_STAGE / DF_BLK) {
    const int lim = min(nkb, base + DF_STAGE / DF_BLK);
    __syncthreads();
    for (int i = base * DF_BLK + tid; i < lim * DF_BLK; i += blockDim.x) stage[i - base * DF_BLK] = (i < k) ? __ldcg(yb + i) : 0.0;
    __syncthreads();
    for (int c = base; c < lim; ++c) {
      double lt[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
      if (c + 1 < nkb) {
        const int t0n = (c + 1) * DF_BLK, ncn = min(DF_BLK, k - t0n);
        const double* nx = P + (size_t)t0n * f;
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr && t < ncn) ? nx[(size_t)t * f] : 0.0; }
      }
      const double* yc = stage + (c - base) * DF_BLK + ty;
#pragma unroll
      for (int q = 0; q < 16; ++q) acc = fma(lt[q], yc[4 * q], acc);
    }
  }
  part[ty * 64 + tx] = acc;
  __syncthreads();
  if (tid < nr) cbv[S.rows_ptr[s] + (rbase - k) + tid] = wmine - (part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid]);
  __syncthreads();
  if (tid == 0) {
    const int nblk = nkb + (f - k + DF_BLK - 1) / DF_BLK;
    const int old = atomicAdd(V.bcnt + s, 1);
    if ((old + 1) % nblk == 0) { st_release(V.done_f + s, epoch); }
  }
}

// sum the per-thread partials pacc[q] (column t = ty + 4q, row lane tx) over the 64 row lanes -> colacc[64] (smem)
__device__ __forceinline__ void reduce_cols(double (&pacc)[16], double* red, double* colacc) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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
}

// backward phase 1, column block b: t_b = z_b - L21[:, b]^T x(contribution rows)   (no dependence on other blocks)
// smem: stage[DF_STAGE] | red[128] | colacc[64]
__device__ void big_bwd_t(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int b, int epoch,
                          double* sm, const double* __restrict__ x) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* red = sm + DF_STAGE;
  double* colacc = red + 128;
  const int t0 = b * DF_BLK, ncol = min(DF_BLK, k - t0);
  const double* __restrict__ P = N.L + S.L_off[s] + (size_t)t0 * f;
  const int nchunk = (r + DF_BLK - 1) / DF_BLK;
  double pacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) pacc[q] = 0.0;
  double ltn[16];
  if (nchunk > 0) {
    const int nr = min(DF_BLK, r);
    const double* base = P + (k + tx);
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr && t < ncol) ? base[(size_t)t * f] : 0.0; }
  }
  const int par = S.sn_parent[s];
  if (par >= 0 && tid == 0) wait_eq(V.done_b + par, epoch);
  __syncthreads();
  for (int cb = 0; cb < nchunk; cb += DF_STAGE / DF_BLK) {
    const int lim = min(nchunk, cb + DF_STAGE / DF_BLK);
    __syncthreads();
    for (int i = cb * DF_BLK + tid; i < lim * DF_BLK; i += blockDim.x) stage[i - cb * DF_BLK] = (i < r) ? __ldcg(x + S.rows[ro + i]) : 0.0;
    __syncthreads();
    for (int ch = cb; ch < lim; ++ch) {
      double lt[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
      if (ch + 1 < nchunk) {
        const int rb2 = k + (ch + 1) * DF_BLK, nr2 = min(DF_BLK, f - rb2);
        const double* base = P + (rb2 + tx);
#pragma unroll
        for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr2 && t < ncol) ? base[(size_t)t * f] : 0.0; }
      }
      const double xi = stage[(ch - cb) * DF_BLK + tx];
#pragma unroll
      for (int q = 0; q < 16; ++q) pacc[q] = fma(lt[q], xi, pacc[q]);
    }
  }
  reduce_cols(pacc, red, colacc);
  if (tid < ncol) V.bigv[V.bigv_off[s] + t0 + tid] = x[c0 + t0 + tid] - colacc[tid];
  __syncthreads();
  if (tid == 0) st_release(V.bflag_b + V.boff[s] + b, epoch);
}

// backward phase 2, column block b: x_b = sum_{c>=b} Linv[c,b]^T t_c
__device__ void big_bwd_x(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int b, int epoch,
                          double* sm, double* __restrict__ x) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int nkb = (k + DF_BLK - 1) / DF_BLK;
  const long long K64 = (long long)nkb * DF_BLK;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* red = sm + DF_STAGE;
  double* colacc = red + 128;
  const int t0 = b * DF_BLK, ncol = min(DF_BLK, k - t0);
  const double* __restrict__ Li = V.linv + V.linv_off[s] + (long long)t0 * K64 + tx;
  const double* __restrict__ tv = V.bigv + V.bigv_off[s];
  double pacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) pacc[q] = 0.0;
  double ltn[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) ltn[q] = Li[(long long)b * DF_BLK + (long long)(ty + 4 * q) * K64];
  wait_flags(V.bflag_b + V.boff[s], b, nkb, epoch);
  for (int cb = b; cb < nkb; cb += DF_STAGE / DF_BLK) {
    const int lim = min(nkb, cb + DF_STAGE / DF_BLK);
    __syncthreads();
    for (int i = cb * DF_BLK + tid; i < lim * DF_BLK; i += blockDim.x) stage[i - cb * DF_BLK] = (i < k) ? __ldcg(tv + i) : 0.0;
    __syncthreads();
    for (int c = cb; c < lim; ++c) {
      double lt[16];
#pragma unroll
      for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
      if (c + 1 < nkb) {
#pragma unroll
        for (int q = 0; q < 16; ++q) ltn[q] = Li[(long long)(c + 1) * DF_BLK + (long long)(ty + 4 * q) * K64];
      }
      const double ti = stage[(c - cb) * DF_BLK + tx];
#pragma unroll
      for (int q = 0; q < 16; ++q) pacc[q] = fma(lt[q], ti, pacc[q]);
    }
  }
  reduce_cols(pacc, red, colacc);
  if (tid < ncol) x[c0 + N.lperm[c0 + t0 + tid]] = colacc[tid];   // final value (the permutation is panel-local)
  __syncthreads();
  if (tid == 0) {
    const int old = atomicAdd(V.bcnt_b + s, 1);
    if ((old + 1) % nkb == 0) { st_release(V.done_b + s, epoch); }
  }
}

// ------------------------------------------------------------------------------------------------
// Explicit inverse of the pivot block L11 of the big fronts, once per factorisation, by recursive doubling:
//   level 0 : the 64x64 diagonal blocks (k_linv_diag: one thread per column, the column in registers);
//   level l : [A 0; B C]^-1 = [A^-1 0; -C^-1 B A^-1  C^-1] for all pairs of neighbouring blocks of 64*2^(l-1)
//             columns, as two batched 64x64-tile GEMM passes (k_linv_gemm<1>: T = B A^-1 into the W scratch of the
//             factorisation, k_linv_gemm<2>: -C^-1 T into Linv).  Work items are enumerated on the host at analysis.
// ------------------------------------------------------------------------------------------------
struct LinvItem { int s, ib, jb, m0, m1; };   // front, tile row / column (64-blocks), k-range of tiles [m0, m1)

#define LI_LD 66
__global__ void __launch_bounds__(64) k_linv_diag(DevSym S, DevNum N, const int* __restrict__ pairs,
                                                  const long long* __restrict__ linv_off, double* __restrict__ linv) {
  __shared__ __align__(16) double Ls[64 * LI_LD];
  const int s = pairs[2 * blockIdx.x], b = pairs[2 * blockIdx.x + 1];
  const int k = S.sn_start[s + 1] - S.sn_start[s];
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const long long K64 = (long long)((k + 63) / 64) * 64;
  const int r0 = b * 64, nd = min(64, k - r0);
  const double* __restrict__ P = N.L + S.L_off[s];
  const int tid = threadIdx.x;
  for (int t = tid; t < 64 * 64; t += 64) {
    const int i = t & 63, q = t >> 6;
    Ls[i + q * LI_LD] = (i < nd && q < nd && i > q) ? P[(r0 + i) + (size_t)(r0 + q) * f] : 0.0;
  }
  __syncthreads();
  // thread j: column j of X = L_dd^-1 by forward substitution on e_j, right-looking (independent FMAs per step)
  double xc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) xc[i] = (i == tid) ? 1.0 : 0.0;
#pragma unroll
  for (int t = 0; t < 63; ++t) {
    const double xt = xc[t];
    const double* __restrict__ col = Ls + t * LI_LD;
    if (((t + 1) & 1) != 0) xc[t + 1] = fma(-xt, col[t + 1], xc[t + 1]);
#pragma unroll
    for (int q = (t + 2) & ~1; q < 64; q += 2) {
      const double2 l2 = *reinterpret_cast<const double2*>(col + q);
      xc[q] = fma(-xt, l2.x, xc[q]);
      xc[q + 1] = fma(-xt, l2.y, xc[q + 1]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 64; ++i) Ls[i + tid * LI_LD] = (i < nd && tid < nd) ? xc[i] : 0.0;
  __syncthreads();
  double* __restrict__ E = linv + linv_off[s] + r0 + (long long)r0 * K64;
  for (int t = tid; t < 64 * 64; t += 64) {
    const int i = t & 63, q = t >> 6;
    E[i + (long long)q * K64] = Ls[i + q * LI_LD];
  }
}

// 64x64 tile per CTA of 128 threads, 8x4 register blocking (rows tx+8q, columns ty+16p), k-slabs of 16 prefetched
// into registers while the current slab is consumed.
template <int PHASE>
__global__ void __launch_bounds__(128) k_linv_gemm(DevSym S, DevNum N, const LinvItem* __restrict__ items,
                                                   const long long* __restrict__ linv_off, double* __restrict__ linv) {
  __shared__ double As[16][65];
  __shared__ double Bs[16][65];
  const LinvItem it = items[blockIdx.x];
  const int s = it.s;
  const int k = S.sn_start[s + 1] - S.sn_start[s];
  const long long f = k + (S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const long long K64 = (long long)((k + 63) / 64) * 64;
  double* __restrict__ Li = linv + linv_off[s];
  const double* __restrict__ Lp = N.L + S.L_off[s];
  double* __restrict__ Wp = N.W + S.L_off[s];
  const int tid = threadIdx.x, tx = tid & 7, ty = tid >> 3;
  const int i0 = it.ib * 64, j0 = it.jb * 64;
  // loaders: A element (row i0 + (tid & 63), col kk0 + (tid >> 6) + 2q); B element (row kk0 + (tid & 15), col j0 + (tid >> 4) + 8q)
  double av[8], bv[8];
  auto load = [&](int kk0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const long long arow = i0 + (tid & 63), acol = kk0 + (tid >> 6) + 2 * q;
      const long long brow = kk0 + (tid & 15), bcol = j0 + (tid >> 4) + 8 * q;
      if (PHASE == 1) {
        av[q] = (arow < k && acol < k) ? Lp[arow + acol * f] : 0.0;
        bv[q] = Li[brow + bcol * K64];
      } else {
        av[q] = Li[arow + acol * K64];
        bv[q] = (brow < k && bcol < k) ? Wp[brow + bcol * f] : 0.0;
      }
    }
  };
  double acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[a][c] = 0.0;
  const int kbeg = it.m0 * 64, kend = it.m1 * 64;
  load(kbeg);
  for (int kk0 = kbeg; kk0 < kend; kk0 += 16) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      As[(tid >> 6) + 2 * q][tid & 63] = av[q];
      Bs[tid & 15][(tid >> 4) + 8 * q] = bv[q];
    }
    __syncthreads();
    if (kk0 + 16 < kend) load(kk0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double a[8], c[4];
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = As[kk][tx + 8 * q];
#pragma unroll
      for (int p = 0; p < 4; ++p) c[p] = Bs[kk][ty + 16 * p];
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int p = 0; p < 4; ++p) acc[q][p] = fma(a[q], c[p], acc[q][p]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 8; ++q)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const long long row = i0 + tx + 8 * q, col = j0 + ty + 16 * p;
      if (PHASE == 1) { if (row < k && col < k) Wp[row + col * f] = acc[q][p]; }
      else Li[row + col * K64] = -acc[q][p];
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
      big_gather(S, N, V, T.a, epoch, sm, x, cbv);
    } else {
      if (FWD) { if (T.c == 0) big_fwd_piv(S, N, V, T.a, T.b, epoch, sm, x); else big_fwd_cb(S, N, V, T.a, T.b, epoch, sm, cbv); }
      else { if (T.c == 0) big_bwd_t(S, N, V, T.a, T.b, epoch, sm, x); else big_bwd_x(S, N, V, T.a, T.b, epoch, sm, x); }
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
