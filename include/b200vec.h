/* b200vec -- C ABI of the device-resident dense-vector / expansion-matrix / triplet-SpMV kernels that replace the
 * host loops of Ipopt's src/LinAlg on the interior-point hot path:
 *   V1-V7  DenseVector::*Impl          reference src/LinAlg/IpDenseVector.cpp:93-1500, IpBlas.cpp:270-298
 *   V8     ExpansionMatrix::*Impl      reference src/LinAlg/IpExpansionMatrix.cpp:27-372
 *   V9     GenTMatrix / SymTMatrix     reference src/LinAlg/TMatrices/IpGenTMatrix.cpp:46-130, IpSymTMatrix.cpp:46-110
 *
 * A vector is described exactly as DenseVector stores it (IpDenseVector.hpp:483-493): `n` entries at `d` (DEVICE
 * memory, capacity n -- the caller always provides it, like values_allocated()), or -- homogeneous != 0 -- the single
 * value `scalar` for all entries.  Every entry point reproduces the reference's homogeneous fast paths (they decide
 * the REPRESENTATION of the result), so a caller that mirrors DenseVector's fields gets the same state machine.
 * Element-wise results are bit-identical to the reference's loops (no FMA contraction: explicit round-to-nearest
 * multiplies/adds/divides in the reference's operation order); min/max-type reductions are exact; sum-type reductions
 * are two-stage warp-shuffle trees with a fixed order (bit-reproducible, not bit-identical to a sequential/BLAS sum).
 * Plain C types only.  All calls on one context are enqueued on its stream; reductions synchronise that stream and
 * return the value on the host (the reference needs it there: IpVector.hpp:429-453 caches it per tag).
 * There is NO CPU fallback: b200vec_create returns NULL when no CUDA device is usable.
 */
#ifndef B200VEC_H
#define B200VEC_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200vec_ctx_s* b200vec_ctx;

typedef struct b200vec {
  double* d;        /* device pointer, capacity n doubles (may be NULL only if n == 0) */
  int n;            /* Dim() */
  int homogeneous;  /* != 0: all entries equal `scalar`, d's contents are unspecified */
  double scalar;
} b200vec;

b200vec_ctx b200vec_create(int device /* -1 = current */, void* stream /* cudaStream_t or NULL */);
void b200vec_destroy(b200vec_ctx c);
const char* b200vec_last_error(b200vec_ctx c);
/* wait for everything enqueued on the context's stream (element-wise calls return without synchronising) */
int b200vec_sync(b200vec_ctx c);
/* kernel launches enqueued since creation (bench bookkeeping) */
int64_t b200vec_launches(b200vec_ctx c);

/* ---- V3: CopyImpl :93-113, ScalImpl :115-128, SetImpl :263-276, AddScalarImpl :618-631 ---- */
int b200vec_copy(b200vec_ctx c, const b200vec* x, b200vec* y);          /* y = x */
int b200vec_scal(b200vec_ctx c, double alpha, b200vec* y);              /* y *= alpha */
int b200vec_set(b200vec_ctx c, double alpha, b200vec* y);               /* y = alpha (homogeneous) */
int b200vec_add_scalar(b200vec_ctx c, double scalar, b200vec* y);       /* y += scalar */
/* ---- V1: AxpyImpl :130-177 (IpBlasAxpy, IpBlas.cpp:270-298) ---- */
int b200vec_axpy(b200vec_ctx c, double alpha, const b200vec* x, b200vec* y);   /* y += alpha x */
/* ---- V2: reductions (DotImpl :179-219, Nrm2 :221-232, Asum :234-245, Amax :247-261, Max/Min :633-679,
 *          Sum/SumLogs :681-721); value returned on the host ---- */
int b200vec_dot(b200vec_ctx c, const b200vec* x, const b200vec* y, double* out);
int b200vec_nrm2(b200vec_ctx c, const b200vec* x, double* out);
int b200vec_asum(b200vec_ctx c, const b200vec* x, double* out);
int b200vec_amax(b200vec_ctx c, const b200vec* x, double* out);
int b200vec_max(b200vec_ctx c, const b200vec* x, double* out);
int b200vec_min(b200vec_ctx c, const b200vec* x, double* out);
int b200vec_sum(b200vec_ctx c, const b200vec* x, double* out);
int b200vec_sumlogs(b200vec_ctx c, const b200vec* x, double* out);
/* ---- V4: element-wise maps (:278-616, 723-759) ---- */
int b200vec_ew_divide(b200vec_ctx c, const b200vec* x, b200vec* y);     /* y /= x */
int b200vec_ew_multiply(b200vec_ctx c, const b200vec* x, b200vec* y);   /* y *= x */
int b200vec_ew_select(b200vec_ctx c, const b200vec* x, b200vec* y);     /* y = y>0 ? x : (y<0 ? -x : 0) */
int b200vec_ew_max(b200vec_ctx c, const b200vec* x, b200vec* y);        /* y = max(y, x) */
int b200vec_ew_min(b200vec_ctx c, const b200vec* x, b200vec* y);        /* y = min(y, x) */
int b200vec_ew_reciprocal(b200vec_ctx c, b200vec* y);
int b200vec_ew_abs(b200vec_ctx c, b200vec* y);
int b200vec_ew_sqrt(b200vec_ctx c, b200vec* y);
int b200vec_ew_sgn(b200vec_ctx c, b200vec* y);
/* ---- V5: AddTwoVectorsImpl :762-1322   y = a v1 + b v2 + c y ---- */
int b200vec_add_two_vectors(b200vec_ctx c, double a, const b200vec* v1, double b, const b200vec* v2, double cc, b200vec* y);
/* ---- V6: FracToBoundImpl :1324-1389    alpha = min(1, min_{delta_i<0} -tau/delta_i * x_i) ---- */
int b200vec_frac_to_bound(b200vec_ctx c, const b200vec* x, const b200vec* delta, double tau, double* out);
/* ---- V7: AddVectorQuotientImpl :1391-1500   y = a z/s + c y ---- */
int b200vec_add_vector_quotient(b200vec_ctx c, double a, const b200vec* z, const b200vec* s, double cc, b200vec* y);

/* ---- V8: ExpansionMatrix (exp_pos: NCols() 0-based row positions in DEVICE memory, ExpandedPosIndices()) ---- */
/* MultVectorImpl :27-96:       y = beta y + alpha P x      (x: ncols, y: nrows) */
int b200vec_exp_mult(b200vec_ctx c, int nrows, int ncols, const int* exp_pos, double alpha, const b200vec* x, double beta, b200vec* y);
/* TransMultVectorImpl :98-167: y = beta y + alpha P^T x    (x: nrows, y: ncols) */
int b200vec_exp_transmult(b200vec_ctx c, int nrows, int ncols, const int* exp_pos, double alpha, const b200vec* x, double beta, b200vec* y);
/* AddMSinvZImpl :170-239:      X += alpha P (Z / S)        (S, Z: ncols, X: nrows; S must be non-homogeneous) */
int b200vec_exp_add_msinvz(b200vec_ctx c, int nrows, int ncols, const int* exp_pos, double alpha, const b200vec* S, const b200vec* Z, b200vec* X);
/* SinvBlrmZMTdBrImpl :241-372: X = (R + alpha Z .* (P^T D)) / S   (S, R, Z, X: ncols, D: nrows; S, D non-homogeneous) */
int b200vec_exp_sinv_blrm_zmtdbr(b200vec_ctx c, int nrows, int ncols, const int* exp_pos, double alpha, const b200vec* S,
                                 const b200vec* R, const b200vec* Z, const b200vec* D, b200vec* X);

/* ---- V9: triplet (COO) matrices.  The structure (1-based irow/jcol, HOST arrays, as GenTMatrixSpace / SymTMatrixSpace
 * hold them) is analysed once: entries are grouped by output row keeping the triplet order, so each y_i is accumulated
 * in exactly the order of the reference's scalar loop (bit-identical, no atomics). ---- */
typedef struct b200vec_tmat_s* b200vec_tmat;
b200vec_tmat b200vec_tmat_create(b200vec_ctx c, int nrows, int ncols, int nnz, const int* irow, const int* jcol, int symmetric);
void b200vec_tmat_destroy(b200vec_tmat m);
/* values: nnz doubles in DEVICE memory (SetValues) */
/* GenTMatrix::MultVectorImpl :46-100 / SymTMatrix::MultVectorImpl :46-110:  y = beta y + alpha A x */
int b200vec_tmat_mult(b200vec_tmat m, const double* values, double alpha, const b200vec* x, double beta, b200vec* y);
/* GenTMatrix::TransMultVectorImpl :102-130:  y = beta y + alpha A^T x  (for a symmetric matrix the same as mult) */
int b200vec_tmat_transmult(b200vec_tmat m, const double* values, double alpha, const b200vec* x, double beta, b200vec* y);

#ifdef __cplusplus
}
#endif
#endif
