// Supernodal triangular solve (forward L, D^-1, backward L^T): TWO kernels per sweep.
//
//   k_solve_sub<FWD> : the BOTTOM of the elimination tree.  The tree below the big separator fronts is cut into
//                      subtrees of bounded size (bytes of L, number of fronts, front order <= 256); ONE CTA walks one
//                      subtree level by level -- fronts of order <= 64 one warp each, larger ones by the whole CTA --
//                      with CTA barriers only: no global flags, no tickets, no atomics.  Subtrees are launched largest
//                      first (the hardware block scheduler then does LPT scheduling).
//   k_solve_top<FWD> : everything above the subtrees: a persistent task-queue kernel.  Fronts up to order 256 are one
//                      task each; a front larger than that is cut into (64-row block) x (4-tile chunk) GEMV tasks on the
//                      EXPLICIT inverse of its pivot block (k_linv_*), so a separator front of order 1000+ keeps ~50
//                      CTAs busy instead of a chain of block steps.  A task waits for its producers through
//                      ld.acquire/st.release flags; it only ever waits for tasks EARLIER in the (topologically sorted)
//                      list and every ticket holder is resident, so the scheme cannot deadlock.  Chunk partials are
//                      combined by the last-arriving CTA in chunk order (fence + counter), so results do not depend on
//                      arrival order: bit-reproducible.
// Forward: sub then top; backward: top then sub (the kernel boundary is the only synchronisation between the two).
// L is streamed exactly once per sweep (HBM-bound, SURVEY.md 8d: 2*8*nnz(L) bytes per right-hand side); children ->
// parent data flows through per-front update vectors gathered by the parent (no atomics on the data path).
// Replaces the vendor back-solve of the reference (MUMPS job=3,
// reference src/Algorithm/LinearSolvers/IpMumpsSolverInterface.cpp:543-583).
#include <cuda_runtime.h>
#include <math.h>

#include "kernels.cuh"

namespace b200 {

#define DF_THREADS 256
#define DF_BLK 64            // block-row / block-column size of the big-front tasks
#define DF_CH 4              // 64x64 tiles per chunk task
#define DF_MIDMAX 256        // fronts above this order are "big" (block tasks + explicit L11 inverse)
#define DF_SMEM_DOUBLES 1600 // max(mid front: 2*256 + 32*33 = 1568, big task: 256 + 256 + 64 + 128 + 64 = 768)

enum { ST_SMALL = 0, ST_MID = 1, ST_FP = 2, ST_FC = 3, ST_BT = 4, ST_BX = 5 };

// ST_SMALL : s = offset into bundle[], blk = number of fronts (<= 8, one warp each, order <= 64)
// ST_MID   : s = front
// ST_FP/FC/BT/BX : s = front, blk = 64-block, tiles [t0, t1) of the contraction, chunk q of nq,
//                  pbase = first partial slot of (s, kind, blk), cidx = its arrival counter
struct SolveTask { int type, s, blk, t0, t1, q, nq, pbase, cidx, pad; };

struct DevSolve {
  const SolveTask* tasks;      // forward list (top part)
  const SolveTask* tasks_bwd;  // backward list
  int ntasks_fwd, ntasks_bwd;
  const int* bundle;           // front ids of the small bundles
  int* done_f;                 // nsn : epoch when the forward work of a front is complete
  int* done_b;                 // nsn : same for backward
  int* bflag_f;                // per (big front, pivot block): y block published
  int* bflag_b;                // per (big front, pivot block): t block published
  int* bcnt;                   // nsn : finished contribution-row blocks (monotonic)
  int* bcnt_b;                 // nsn : finished backward x blocks (monotonic)
  const int* boff;             // nsn : offset of a big front's blocks in bflag_*
  const long long* bigv_off;   // nsn : offset into bigv / bigy (f doubles per big front)
  double* bigv;                // big fronts: z = D^-1 y (forward), then t (backward), pivoted order
  double* bigy;                // big fronts: y (forward), pivoted order
  double* part;                // chunk partials, 64 doubles per slot
  int* ccnt;                   // chunk arrival counters (monotonic, modulo nq)
  unsigned long long* ticket;  // [0] fwd, [1] bwd (monotonic)
  const double* linv;          // explicit inverses of the big fronts' pivot blocks L11 (K64 x K64 each, see k_linv_*)
  const long long* linv_off;   // nsn : offset of a big front's inverse in linv, -1 = none
  // subtrees (k_solve_sub): subtree u owns levels [sub_ptr[u], sub_ptr[u+1]) of lvl_*; level e owns the fronts
  // sub_fronts[lvl_ptr[e] .. lvl_ptr[e+1]) -- the first lvl_nsmall[e] of them have order <= 64
  const int* sub_ptr;
  const int* lvl_ptr;
  const int* lvl_nsmall;
  const int* sub_fronts;
  const int* sub_root;         // root front of each subtree (its done_f is published for the top kernel)
  int nsub;
  unsigned long long* tlog;    // optional (debug): 2 timestamps per top task, fwd then bwd; nullptr = off
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
  while (ld_acquire(p) != epoch) __nanosleep(20);
}

// ------------------------------------------------------------------------------------------------
// fronts of order <= 64: one warp per front, two rows per lane (lane, lane+32), L streamed straight from global
// memory: the loads do not depend on the running vector, so 16 columns are always in flight ahead of the shuffle chain.
// FLAGS = true : top kernel (wait for the children / the parent, publish done flags)
// FLAGS = false: subtree kernel (ordering comes from CTA barriers)
// smem per warp: w[64]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void w64_load8(const double* __restrict__ P, int f, int k, int i0, int i1, int tb,
                                          double (&l0)[8], double (&l1)[8]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int t = tb + q;
    l0[q] = (t < k && i0 > t && i0 < f) ? P[i0 + (size_t)t * f] : 0.0;
    l1[q] = (t < k && i1 > t && i1 < f) ? P[i1 + (size_t)t * f] : 0.0;
  }
}
__device__ __forceinline__ void w64_fstep8(int k, int tb, const double (&l0)[8], const double (&l1)[8], double& v0,
                                           double& v1) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int t = tb + q;
    if (t < k) {   // warp-uniform
      const double yt = (t < 32) ? __shfl_sync(0xffffffffu, v0, t) : __shfl_sync(0xffffffffu, v1, t - 32);
      v0 = fma(-l0[q], yt, v0);
      v1 = fma(-l1[q], yt, v1);
    }
  }
}

template <bool FLAGS>
__device__ void w64_fwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* w,
                        double* __restrict__ x, double* __restrict__ cbv) {
  const int lane = threadIdx.x & 31;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  const int ch0 = S.child_ptr[s], nch = S.child_ptr[s + 1] - ch0;
  const int i0 = lane, i1 = lane + 32;
  const double* __restrict__ P = N.L + S.L_off[s];
  double a0[8], a1[8], b0[8], b1[8];
  w64_load8(P, f, k, i0, i1, 0, a0, a1);            // in flight while the children are gathered
  w[i0] = (i0 < k) ? x[c0 + i0] : 0.0;
  w[i1] = (i1 < k) ? x[c0 + i1] : 0.0;
  __syncwarp();
  for (int q0 = 0; q0 < nch; q0 += 32) {
    // child metadata lane-parallel, then the children one after the other (their targets may overlap); the
    // (index, value) pairs of child q+1 are in flight while child q is added
    const int m = min(32, nch - q0);
    int cq = -1, rq = 0;
    long long oq = 0;
    if (lane < m) {
      cq = S.child_idx[ch0 + q0 + lane];
      oq = S.rows_ptr[cq];
      rq = (int)(S.rows_ptr[cq + 1] - oq);
      if (FLAGS) wait_eq(V.done_f + cq, epoch);
    }
    __syncwarp();
    int nidx0 = 0, nidx1 = 0;
    double nval0 = 0.0, nval1 = 0.0;
    bool nok0 = false, nok1 = false;
    {
      const long long o = __shfl_sync(0xffffffffu, oq, 0);
      const int rc = __shfl_sync(0xffffffffu, rq, 0);
      nok0 = lane < rc; nok1 = lane + 32 < rc;
      if (nok0) { nidx0 = S.rel[o + lane]; nval0 = __ldcg(cbv + o + lane); }
      if (nok1) { nidx1 = S.rel[o + lane + 32]; nval1 = __ldcg(cbv + o + lane + 32); }
    }
    for (int q = 0; q < m; ++q) {
      const int idx0 = nidx0, idx1 = nidx1;
      const double val0 = nval0, val1 = nval1;
      const bool ok0 = nok0, ok1 = nok1;
      if (q + 1 < m) {
        const long long o = __shfl_sync(0xffffffffu, oq, q + 1);
        const int rc = __shfl_sync(0xffffffffu, rq, q + 1);
        nok0 = lane < rc; nok1 = lane + 32 < rc;
        if (nok0) { nidx0 = S.rel[o + lane]; nval0 = __ldcg(cbv + o + lane); }
        if (nok1) { nidx1 = S.rel[o + lane + 32]; nval1 = __ldcg(cbv + o + lane + 32); }
      }
      if (ok0) w[idx0] += val0;   // rel is strictly increasing inside a child: no two lanes hit the same entry
      if (ok1) w[idx1] += val1;
      __syncwarp();
    }
  }
  double v0 = 0.0, v1 = 0.0;
  if (i0 < f) v0 = (i0 < k) ? w[N.lperm[c0 + i0]] : w[i0];
  if (i1 < f) v1 = (i1 < k) ? w[N.lperm[c0 + i1]] : w[i1];
  for (int tb = 0; tb < k; tb += 16) {
    if (tb + 8 < k) w64_load8(P, f, k, i0, i1, tb + 8, b0, b1);
    w64_fstep8(k, tb, a0, a1, v0, v1);
    if (tb + 16 < k) w64_load8(P, f, k, i0, i1, tb + 16, a0, a1);
    if (tb + 8 < k) w64_fstep8(k, tb + 8, b0, b1, v0, v1);
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
  if (FLAGS && lane == 0) st_release(V.done_f + s, epoch);
}

// sum over the 32 lanes of p[t], t = 0..31: lane t receives the total of column t (31 shuffles instead of 32 x 5)
__device__ __forceinline__ double warp_transpose_reduce(double (&p)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const bool hi = lane & 16;
    const double send = hi ? p[j] : p[j + 16], keep = hi ? p[j + 16] : p[j];
    p[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool hi = lane & 8;
    const double send = hi ? p[j] : p[j + 8], keep = hi ? p[j + 8] : p[j];
    p[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool hi = lane & 4;
    const double send = hi ? p[j] : p[j + 4], keep = hi ? p[j + 4] : p[j];
    p[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const bool hi = lane & 2;
    const double send = hi ? p[j] : p[j + 2], keep = hi ? p[j + 2] : p[j];
    p[j] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  {
    const bool hi = lane & 1;
    const double send = hi ? p[0] : p[1], keep = hi ? p[1] : p[0];
    p[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  }
  return p[0];
}

template <bool FLAGS>
__device__ void w64_bwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch,
                        double* __restrict__ x) {
  const int lane = threadIdx.x & 31;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  const int i0 = lane, i1 = lane + 32;
  const double* __restrict__ P = N.L + S.L_off[s];
  if (FLAGS) {
    const int par = S.sn_parent[s];
    if (par >= 0 && lane == 0) wait_eq(V.done_b + par, epoch);
    __syncwarp();
  }
  double v0 = 0.0, v1 = 0.0;  // entries i0 / i1 of [D^-1 y ; x(rows)]
  if (i0 < k) v0 = x[c0 + i0]; else if (i0 < f) v0 = __ldcg(x + S.rows[ro + i0 - k]);
  if (i1 < k) v1 = x[c0 + i1]; else if (i1 < f) v1 = __ldcg(x + S.rows[ro + i1 - k]);
  if (k <= 32) {
    // (1) rectangular part  u_t = sum_{i >= k} L[i,t] v_i : no chain -- coalesced column loads (lane = row), per-lane
    //     products for all 32 columns, one transposing warp reduction (lane t gets u_t)
    const bool cb0 = i0 >= k && i0 < f, cb1 = i1 < f;
    const double m0 = cb0 ? v0 : 0.0, m1 = cb1 ? v1 : 0.0;
    double p[32];
#pragma unroll
    for (int tb = 0; tb < 32; tb += 8) {
      double l0[8], l1[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int t = tb + q;
        l0[q] = (t < k && cb0) ? P[i0 + (size_t)t * f] : 0.0;
        l1[q] = (t < k && cb1) ? P[i1 + (size_t)t * f] : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) p[tb + q] = fma(l0[q], m0, l1[q] * m1);
    }
    // (2) triangle: lane t owns column t; its sub-diagonal entries L[ii, t] (ii > t) are 31 independent loads
    //     (column t is contiguous: 16 consecutive ii share a 128-byte line), then a chain of k-1 shuffle+FMA steps
    double lr[32];
#pragma unroll
    for (int ii = 1; ii < 32; ++ii) lr[ii] = (ii < k && lane < ii) ? P[ii + (size_t)lane * f] : 0.0;
    const double u = warp_transpose_reduce(p);
    double z = (lane < k) ? v0 - u : 0.0;
#pragma unroll
    for (int ii = 31; ii >= 1; --ii) {
      if (ii < k) {   // warp-uniform
        const double xi = __shfl_sync(0xffffffffu, z, ii);
        z = fma(-lr[ii], xi, z);
      }
    }
    if (lane < k) x[c0 + N.lperm[c0 + lane]] = z;
  } else {
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
  }
  __syncwarp();
  if (FLAGS && lane == 0) st_release(V.done_b + s, epoch);
}

// ------------------------------------------------------------------------------------------------
// mid fronts (65 .. 256): one CTA, blocked by 32
// smem: v[f] | w[f] | Lb[32*33]
// ------------------------------------------------------------------------------------------------
template <bool FLAGS>
__device__ void mid_fwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                        double* __restrict__ x, double* __restrict__ cbv) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int r = (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]), f = k + r;
  double* v = sm;
  double* w = sm + f;
  double* Lb = sm + 2 * f;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
  const int ch0 = S.child_ptr[s], ch1 = S.child_ptr[s + 1];
  if (FLAGS) for (int q = ch0 + tid; q < ch1; q += nt) wait_eq(V.done_f + S.child_idx[q], epoch);
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
  if (FLAGS && tid == 0) st_release(V.done_f + s, epoch);
}

template <bool FLAGS>
__device__ void mid_bwd(const DevSym& S, const DevNum& N, const DevSolve& V, int s, int epoch, double* sm,
                        double* __restrict__ x) {
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  double* v = sm;
  double* Lb = sm + 2 * f;
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
  if (FLAGS) {
    const int par = S.sn_parent[s];
    if (par >= 0 && tid == 0) wait_eq(V.done_b + par, epoch);
  }
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
  if (FLAGS && tid == 0) st_release(V.done_b + s, epoch);
}

// ------------------------------------------------------------------------------------------------
// Big fronts (order > 256).  With the EXPLICIT inverse of the unit-lower-triangular pivot block L11 (k_linv_* below,
// once per factorisation) the in-front recurrences become block GEMVs with no chain:
//   forward   FP: y_b   = sum_{c<=b} Linv[b,c] w_c          (w = assembled, pivot-permuted right-hand side)
//             FC: u_j   = w_j - L21[j,:] y                   (update vector handed to the parent)
//   backward  BT: t_b   = z_b - L21[:,b]^T x(rows)           (z = D^-1 y)
//             BX: x_b   = sum_{c>=b} Linv[c,b]^T t_c
// Each (64-block, chunk of <= DF_CH tiles) is one task; the w entries a task needs are gathered on the fly from the
// children's update vectors through the inverse row maps (S.einv) -- no separate gather task, no staging buffer.
// The right-hand side of a big front stays in x[c0..c0+k) until BX overwrites it with the solution: the forward result
// z lives in V.bigv (so FP tasks of other blocks can still read the right-hand side).
// smem: stage[256] | part[256] | ys[64] | red[128] | colacc[64]
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double big_gather_row(const DevSym& S, int j, int k, int c0, int ch0, int nch,
                                                 const double* __restrict__ x, const double* __restrict__ cbv) {
  double v = (j < k) ? x[c0 + j] : 0.0;
  for (int q = 0; q < nch; ++q) {   // fixed child order: deterministic
    const int c = S.child_idx[ch0 + q];
    const int e = S.einv[S.einv_off[c] + j];
    if (e >= 0) v += __ldcg(cbv + S.rows_ptr[c] + e);
  }
  return v;
}

// Combine the chunk partials of one (front, kind, block): every CTA stores its 64 partial sums, the LAST one to arrive
// adds them in chunk order (independent of the arrival order).  `out` (shared, 64 doubles) holds this CTA's partial on
// entry and the total on exit.  Returns true (in all threads) for the CTA that must finalise the block.
__device__ __forceinline__ bool chunk_combine(const DevSolve& V, const SolveTask& T, double* out, int* s_flag) {
  if (T.nq == 1) return true;
  const int tid = threadIdx.x;
  if (tid < 64) __stcg(V.part + ((long long)T.pbase + T.q) * 64 + tid, out[tid]);
  __syncthreads();
  if (tid == 0) {
    __threadfence();                                    // release: this CTA's partial before the counter
    const int old = atomicAdd(V.ccnt + T.cidx, 1);
    const int last = ((old + 1) % T.nq) == 0;
    if (last) __threadfence();                          // acquire: the other CTAs' partials after the counter
    *s_flag = last;
  }
  __syncthreads();
  if (!*s_flag) return false;
  if (tid < 64) {
    const double* base = V.part + (long long)T.pbase * 64 + tid;
    double a = 0.0;
    for (int q = 0; q < T.nq; ++q) a += __ldcg(base + (long long)q * 64);
    out[tid] = a;
  }
  __syncthreads();
  return true;
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

// FP: pivot rows [64 blk, +64) of front s, tiles [t0, t1) of Linv's block row
__device__ void big_fp(const DevSym& S, const DevNum& N, const DevSolve& V, const SolveTask& T, int epoch, double* sm,
                       int* s_flag, const double* __restrict__ x, const double* __restrict__ cbv) {
  const int s = T.s, rb = T.blk;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int nkb = (k + DF_BLK - 1) / DF_BLK;
  const long long K64 = (long long)nkb * DF_BLK;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* part = sm + 256;
  double* ys = part + 256;
  const int r0 = rb * DF_BLK, nrow = min(DF_BLK, k - r0);
  const double* __restrict__ Li = V.linv + V.linv_off[s] + (r0 + tx);
  // first tile in flight before the children are seen (Linv does not depend on the right-hand side)
  double ltn[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) ltn[q] = Li[((long long)T.t0 * DF_BLK + ty + 4 * q) * K64];
  const int ch0 = S.child_ptr[s], nch = S.child_ptr[s + 1] - ch0;
  for (int q = tid; q < nch; q += DF_THREADS) wait_eq(V.done_f + S.child_idx[ch0 + q], epoch);
  __syncthreads();
  {
    const int* __restrict__ lp = N.lperm + c0;
    const int ncols = (T.t1 - T.t0) * DF_BLK;
    for (int i = tid; i < ncols; i += DF_THREADS) {
      const int gi = T.t0 * DF_BLK + i;
      stage[i] = (gi < k) ? big_gather_row(S, lp[gi], k, c0, ch0, nch, x, cbv) : 0.0;
    }
  }
  __syncthreads();
  double acc = 0.0;
  for (int c = T.t0; c < T.t1; ++c) {
    double lt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
    if (c + 1 < T.t1) {
      const double* nx = Li + (long long)(c + 1) * DF_BLK * K64;
#pragma unroll
      for (int q = 0; q < 16; ++q) ltn[q] = nx[(long long)(ty + 4 * q) * K64];
    }
    const double* wc = stage + (c - T.t0) * DF_BLK + ty;
#pragma unroll
    for (int q = 0; q < 16; ++q) acc = fma(lt[q], wc[4 * q], acc);
  }
  part[ty * 64 + tx] = acc;
  __syncthreads();
  if (tid < 64) ys[tid] = part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid];
  __syncthreads();
  if (!chunk_combine(V, T, ys, s_flag)) return;
  if (tid < nrow) {
    const int i = r0 + tid;
    const long long o = V.bigv_off[s];
    V.bigy[o + i] = ys[tid];              // y block for the contribution rows (pivoted order)
    const int ty2 = N.ptype[c0 + i];      // 2x2 partners never straddle a 32-column panel, so they sit in this block
    double z;
    if (ty2 == 1) z = ys[tid] * N.dinv[c0 + i];
    else if (ty2 == 2) z = ys[tid] * N.dinv[c0 + i] + ys[tid + 1] * N.doff[c0 + i];
    else z = ys[tid - 1] * N.doff[c0 + i - 1] + ys[tid] * N.dinv[c0 + i];
    V.bigv[o + i] = z;
  }
  __syncthreads();
  if (tid == 0) st_release(V.bflag_f + V.boff[s] + rb, epoch);
}

// FC: contribution rows [k + 64 blk, +64), tiles [t0, t1) of L21's block row
__device__ void big_fc(const DevSym& S, const DevNum& N, const DevSolve& V, const SolveTask& T, int epoch, double* sm,
                       int* s_flag, const double* __restrict__ x, double* __restrict__ cbv) {
  const int s = T.s, j = T.blk;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int f = k + (int)(S.rows_ptr[s + 1] - S.rows_ptr[s]);
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* part = sm + 256;
  double* ys = part + 256;
  const int rbase = k + j * DF_BLK, nr = min(DF_BLK, f - rbase);
  const double* __restrict__ P = N.L + S.L_off[s] + (rbase + tx);
  double ltn[16];
  {
    const int tc0 = T.t0 * DF_BLK, ncn = min(DF_BLK, k - tc0);
    const double* nx = P + (size_t)tc0 * f;
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr && t < ncn) ? nx[(size_t)t * f] : 0.0; }
  }
  for (int c = T.t0 + tid; c < T.t1; c += DF_THREADS) wait_eq(V.bflag_f + V.boff[s] + c, epoch);
  __syncthreads();
  {
    const double* __restrict__ yb = V.bigy + V.bigv_off[s];
    const int ncols = (T.t1 - T.t0) * DF_BLK;
    for (int i = tid; i < ncols; i += DF_THREADS) {
      const int gi = T.t0 * DF_BLK + i;
      stage[i] = (gi < k) ? __ldcg(yb + gi) : 0.0;
    }
  }
  __syncthreads();
  double acc = 0.0;
  for (int c = T.t0; c < T.t1; ++c) {
    double lt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
    if (c + 1 < T.t1) {
      const int tcn = (c + 1) * DF_BLK, ncn = min(DF_BLK, k - tcn);
      const double* nx = P + (size_t)tcn * f;
#pragma unroll
      for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr && t < ncn) ? nx[(size_t)t * f] : 0.0; }
    }
    const double* yc = stage + (c - T.t0) * DF_BLK + ty;
#pragma unroll
    for (int q = 0; q < 16; ++q) acc = fma(lt[q], yc[4 * q], acc);
  }
  part[ty * 64 + tx] = acc;
  __syncthreads();
  if (tid < 64) ys[tid] = part[tid] + part[64 + tid] + part[128 + tid] + part[192 + tid];
  __syncthreads();
  if (!chunk_combine(V, T, ys, s_flag)) return;
  if (tid < nr) {
    // the children are complete (every FP task of this front waited for them before publishing the flags seen above)
    const int ch0 = S.child_ptr[s], nch = S.child_ptr[s + 1] - ch0;
    const double wv = big_gather_row(S, rbase + tid, k, c0, ch0, nch, x, cbv);
    cbv[S.rows_ptr[s] + (rbase - k) + tid] = wv - ys[tid];
  }
  __syncthreads();
  if (tid == 0) {
    const int ncb = (f - k + DF_BLK - 1) / DF_BLK;
    __threadfence();                                 // this block's update vector before the counter
    const int old = atomicAdd(V.bcnt + s, 1);
    if ((old + 1) % ncb == 0) { __threadfence(); st_release(V.done_f + s, epoch); }
  }
}

// BT: column block blk, contribution-row tiles [t0, t1):  t_b = z_b - L21[:, b]^T x(rows)
__device__ void big_bt(const DevSym& S, const DevNum& N, const DevSolve& V, const SolveTask& T, int epoch, double* sm,
                       int* s_flag, const double* __restrict__ x) {
  const int s = T.s, b = T.blk;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const long long ro = S.rows_ptr[s];
  const int r = (int)(S.rows_ptr[s + 1] - ro), f = k + r;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* red = sm + 576;
  double* colacc = red + 128;
  const int tc0 = b * DF_BLK, ncol = min(DF_BLK, k - tc0);
  const double* __restrict__ P = N.L + S.L_off[s] + (size_t)tc0 * f;
  double pacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) pacc[q] = 0.0;
  double ltn[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) ltn[q] = 0.0;
  if (T.t1 > T.t0) {
    const int rb2 = k + T.t0 * DF_BLK, nr2 = min(DF_BLK, f - rb2);
    const double* base = P + (rb2 + tx);
#pragma unroll
    for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr2 && t < ncol) ? base[(size_t)t * f] : 0.0; }
  }
  {
    const int par = S.sn_parent[s];
    if (par >= 0 && tid == 0) wait_eq(V.done_b + par, epoch);
  }
  __syncthreads();
  {
    const int nrows = (T.t1 - T.t0) * DF_BLK;
    for (int i = tid; i < nrows; i += DF_THREADS) {
      const int ri = T.t0 * DF_BLK + i;
      stage[i] = (ri < r) ? __ldcg(x + S.rows[ro + ri]) : 0.0;
    }
  }
  __syncthreads();
  for (int ch = T.t0; ch < T.t1; ++ch) {
    double lt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
    if (ch + 1 < T.t1) {
      const int rb2 = k + (ch + 1) * DF_BLK, nr2 = min(DF_BLK, f - rb2);
      const double* base = P + (rb2 + tx);
#pragma unroll
      for (int q = 0; q < 16; ++q) { const int t = ty + 4 * q; ltn[q] = (tx < nr2 && t < ncol) ? base[(size_t)t * f] : 0.0; }
    }
    const double xi = stage[(ch - T.t0) * DF_BLK + tx];
#pragma unroll
    for (int q = 0; q < 16; ++q) pacc[q] = fma(lt[q], xi, pacc[q]);
  }
  reduce_cols(pacc, red, colacc);
  if (!chunk_combine(V, T, colacc, s_flag)) return;
  if (tid < ncol) {
    double* tv = V.bigv + V.bigv_off[s] + tc0 + tid;
    *tv = *tv - colacc[tid];            // z (written by FP in the forward sweep) -> t, in place
  }
  __syncthreads();
  if (tid == 0) st_release(V.bflag_b + V.boff[s] + b, epoch);
}

// BX: column block blk, tiles [t0, t1) of Linv's block column (t0 >= blk):  x_b = sum_c Linv[c,b]^T t_c
__device__ void big_bx(const DevSym& S, const DevNum& N, const DevSolve& V, const SolveTask& T, int epoch, double* sm,
                       int* s_flag, double* __restrict__ x) {
  const int s = T.s, b = T.blk;
  const int c0 = S.sn_start[s], k = S.sn_start[s + 1] - c0;
  const int nkb = (k + DF_BLK - 1) / DF_BLK;
  const long long K64 = (long long)nkb * DF_BLK;
  const int tid = threadIdx.x, tx = tid & 63, ty = tid >> 6;
  double* stage = sm;
  double* red = sm + 576;
  double* colacc = red + 128;
  const int tc0 = b * DF_BLK, ncol = min(DF_BLK, k - tc0);
  const double* __restrict__ Li = V.linv + V.linv_off[s] + (long long)tc0 * K64 + tx;
  const double* __restrict__ tv = V.bigv + V.bigv_off[s];
  double pacc[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) pacc[q] = 0.0;
  double ltn[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) ltn[q] = Li[(long long)T.t0 * DF_BLK + (long long)(ty + 4 * q) * K64];
  for (int c = T.t0 + tid; c < T.t1; c += DF_THREADS) wait_eq(V.bflag_b + V.boff[s] + c, epoch);
  __syncthreads();
  {
    const int nrows = (T.t1 - T.t0) * DF_BLK;
    for (int i = tid; i < nrows; i += DF_THREADS) {
      const int gi = T.t0 * DF_BLK + i;
      stage[i] = (gi < k) ? __ldcg(tv + gi) : 0.0;
    }
  }
  __syncthreads();
  for (int c = T.t0; c < T.t1; ++c) {
    double lt[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) lt[q] = ltn[q];
    if (c + 1 < T.t1) {
#pragma unroll
      for (int q = 0; q < 16; ++q) ltn[q] = Li[(long long)(c + 1) * DF_BLK + (long long)(ty + 4 * q) * K64];
    }
    const double ti = stage[(c - T.t0) * DF_BLK + tx];
#pragma unroll
    for (int q = 0; q < 16; ++q) pacc[q] = fma(lt[q], ti, pacc[q]);
  }
  reduce_cols(pacc, red, colacc);
  if (!chunk_combine(V, T, colacc, s_flag)) return;
  if (tid < ncol) x[c0 + N.lperm[c0 + tc0 + tid]] = colacc[tid];   // final value (the permutation is panel-local)
  __syncthreads();
  if (tid == 0) {
    __threadfence();                                 // this block's solution entries before the counter
    const int old = atomicAdd(V.bcnt_b + s, 1);
    if ((old + 1) % nkb == 0) { __threadfence(); st_release(V.done_b + s, epoch); }
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
// bottom of the tree: one CTA per subtree, level by level, CTA barriers only
// ------------------------------------------------------------------------------------------------
template <bool FWD>
__global__ void __launch_bounds__(DF_THREADS, 2) k_solve_sub(DevSym S, DevNum N, DevSolve V, int epoch,
                                                          double* __restrict__ x, double* __restrict__ cbv) {
  __shared__ double sm[DF_SMEM_DOUBLES];
  const int u = blockIdx.x;
  const int e0 = V.sub_ptr[u], nlv = V.sub_ptr[u + 1] - e0;
  const int warp = threadIdx.x >> 5;
  for (int li = 0; li < nlv; ++li) {
    const int e = e0 + (FWD ? li : nlv - 1 - li);
    const int b = V.lvl_ptr[e], en = V.lvl_ptr[e + 1], ns = V.lvl_nsmall[e];
    for (int q = b + warp; q < b + ns; q += DF_THREADS / 32) {
      const int s = V.sub_fronts[q];
      if (FWD) w64_fwd<false>(S, N, V, s, epoch, sm + warp * 64, x, cbv);
      else w64_bwd<false>(S, N, V, s, epoch, x);
    }
    if (en > b + ns) {
      __syncthreads();   // the per-warp scratch of the small fronts overlaps the mid-front buffers
      for (int q = b + ns; q < en; ++q) {
        const int s = V.sub_fronts[q];
        if (FWD) mid_fwd<false>(S, N, V, s, epoch, sm, x, cbv);
        else mid_bwd<false>(S, N, V, s, epoch, sm, x);
      }
    }
    __syncthreads();     // (also orders this level's global writes before the next level's reads, CTA scope)
  }
  if (FWD && threadIdx.x == 0) V.done_f[V.sub_root[u]] = epoch;   // read by the top kernel (next launch)
}

// ------------------------------------------------------------------------------------------------
// top of the tree: persistent task queue
// ------------------------------------------------------------------------------------------------
template <bool FWD>
__global__ void __launch_bounds__(DF_THREADS, 2) k_solve_top(DevSym S, DevNum N, DevSolve V, int epoch,
                                                          unsigned long long ticket_base,
                                                          double* __restrict__ x, double* __restrict__ cbv) {
  __shared__ double sm[DF_SMEM_DOUBLES];
  __shared__ unsigned long long s_ticket;
  __shared__ int s_flag;
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
    if (V.tlog && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    if (T.type == ST_SMALL) {
      const int w = threadIdx.x >> 5;
      if (w < T.blk) {
        const int s = V.bundle[T.s + w];
        if (FWD) w64_fwd<true>(S, N, V, s, epoch, sm + w * 64, x, cbv);
        else w64_bwd<true>(S, N, V, s, epoch, x);
      }
    } else if (T.type == ST_MID) {
      if (FWD) mid_fwd<true>(S, N, V, T.s, epoch, sm, x, cbv); else mid_bwd<true>(S, N, V, T.s, epoch, sm, x);
    } else if (T.type == ST_FP) {
      big_fp(S, N, V, T, epoch, sm, &s_flag, x, cbv);
    } else if (T.type == ST_FC) {
      big_fc(S, N, V, T, epoch, sm, &s_flag, x, cbv);
    } else if (T.type == ST_BT) {
      big_bt(S, N, V, T, epoch, sm, &s_flag, x);
    } else {
      big_bx(S, N, V, T, epoch, sm, &s_flag, x);
    }
    __syncthreads();
    if (V.tlog && threadIdx.x == 0) {
      unsigned long long t_end;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
      unsigned long long* rec = V.tlog + 2 * ((FWD ? 0 : (unsigned long long)V.ntasks_fwd) + tk);
      rec[0] = t_start; rec[1] = t_end;
    }
  }
}

}  // namespace b200
