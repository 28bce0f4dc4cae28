// Device-side data descriptors shared by the numeric kernels (sm_100a).
#pragma once
#include <cstdint>

namespace b200 {

// counters[] layout
enum { CNT_NEG = 0, CNT_FORCED = 1, CNT_TINY = 2, CNT_GROWTH = 3, CNT_2X2 = 4, CNT_N = 8 };

struct DevSym {
  int n, nsn;
  const int* sn_start;         // nsn+1
  const int* sn_parent;        // nsn
  const long long* rows_ptr;   // nsn+1
  const int* rows;             // permuted row ids
  const int* rel;              // index in the parent front
  const int* child_ptr;        // nsn+1
  const int* child_idx;
  const long long* uent_ptr;   // nsn+1
  const unsigned* u_dst;       // (lcol<<16)|lrow
  const long long* u_dst64;    // lrow + lcol*f
  const long long* L_off;      // nsn+1
  const long long* cb_off;     // nsn+1
  // children of the big fronts: inverse of rel -- einv[einv_off[c] + j] = index of parent-front row j in child c's
  // row list, or -1 (einv_off[c] = -1 if c's parent is not a big front)
  const long long* einv_off;   // nsn
  const int* einv;
};

struct DevNum {
  double* L;        // supernodal panels, f x k column-major, unit diagonal implied
  double* W;        // L*D for big fronts (same offsets as L), scratch
  double* CB;       // contribution blocks r x r (lower part valid)
  double* uval;     // summed+scaled unique entries
  double* dinv;     // n : D^-1 diagonal
  double* doff;     // n : D^-1 off-diagonal at the first column of a 2x2 pivot
  int* ptype;       // n : 1 = 1x1, 2 = first of 2x2, 3 = second of 2x2
  int* lperm;       // n : pivot position t of supernode s holds pre-pivot local column lperm[start+t]
  int* bperm;       // n : block-local permutation of the last panel step (big fronts)
  double* colmax;   // n : big fronts only: max |entry| of a panel column BELOW its 32x32 diagonal block at panel start
  int* counters;    // CNT_N
  double u;         // pivot threshold
  double tiny;      // zero-pivot threshold (scaled matrix)
};

}  // namespace b200
