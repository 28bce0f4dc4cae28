// TEST INFRASTRUCTURE (checker, never shipped or timed as the product): drives the REFERENCE's own classes --
// Ipopt::DenseVector, ExpansionMatrix, GenTMatrix, SymTMatrix from the unmodified libipopt.so built into oracle/_ref --
// on host arrays, behind a tiny C ABI that tests/test_vec_parity.py loads with ctypes.  It is the oracle of the
// b200vec kernels: same inputs through the reference implementation of each *Impl method
// (reference src/LinAlg/IpDenseVector.cpp:93-1500, IpExpansionMatrix.cpp:27-372, TMatrices/IpGenTMatrix.cpp:46-130,
// TMatrices/IpSymTMatrix.cpp:46-110).
#include <cstring>
#include <string>

#include "IpDenseVector.hpp"
#include "IpExpansionMatrix.hpp"
#include "IpGenTMatrix.hpp"
#include "IpSymTMatrix.hpp"
#include "IpCompoundSymMatrix.hpp"
#include "IpDiagMatrix.hpp"
#include "IpIdentityMatrix.hpp"
#include "IpSumSymMatrix.hpp"
#include "IpTripletHelper.hpp"

using namespace Ipopt;

namespace
{
struct VecIn { const double* d; int h; double s; };

SmartPtr<DenseVector> make(const DenseVectorSpace& sp, const double* d, int h, double s)
{
   SmartPtr<DenseVector> v = sp.MakeNewDenseVector();
   if( h ) v->Set(s);
   else if( sp.Dim() > 0 ) v->SetValues(d);
   else v->Set(0.);   // dimension 0: nothing to store
   return v;
}

void read_back(const DenseVector& v, double* d, int* h, double* s)
{
   *h = v.IsHomogeneous() ? 1 : 0;
   if( v.IsHomogeneous() ) *s = v.Scalar();
   else if( v.Dim() > 0 ) memcpy(d, v.Values(), sizeof(double) * v.Dim());
}
}

extern "C"
{

// op on DenseVectors of dimension n.  x1 / x2: inputs (x, v1 / v2, z / s, delta); y: the in/out vector ("this");
// out: value of a reduction.  Returns 0, or 1 for an unknown op.
int vecref_op(const char* opname, int n, double a, double b, double c,
              const double* x1, int h1, double s1, const double* x2, int h2, double s2,
              double* y, int* hy, double* sy, double* out)
{
   const std::string op(opname);
   SmartPtr<DenseVectorSpace> sp = new DenseVectorSpace(n);
   SmartPtr<DenseVector> X1 = make(*sp, x1, h1, s1);
   SmartPtr<DenseVector> X2 = make(*sp, x2, h2, s2);
   SmartPtr<DenseVector> Y = make(*sp, y, *hy, *sy);
   if( op == "copy" ) Y->Copy(*X1);
   else if( op == "scal" ) Y->Scal(a);
   else if( op == "set" ) Y->Set(a);
   else if( op == "add_scalar" ) Y->AddScalar(a);
   else if( op == "axpy" ) Y->Axpy(a, *X1);
   else if( op == "dot" ) *out = Y->Dot(*X1);
   else if( op == "nrm2" ) *out = Y->Nrm2();
   else if( op == "asum" ) *out = Y->Asum();
   else if( op == "amax" ) *out = Y->Amax();
   else if( op == "max" ) *out = Y->Max();
   else if( op == "min" ) *out = Y->Min();
   else if( op == "sum" ) *out = Y->Sum();
   else if( op == "sumlogs" ) *out = Y->SumLogs();
   else if( op == "ew_divide" ) Y->ElementWiseDivide(*X1);
   else if( op == "ew_multiply" ) Y->ElementWiseMultiply(*X1);
   else if( op == "ew_select" ) Y->ElementWiseSelect(*X1);
   else if( op == "ew_max" ) Y->ElementWiseMax(*X1);
   else if( op == "ew_min" ) Y->ElementWiseMin(*X1);
   else if( op == "ew_reciprocal" ) Y->ElementWiseReciprocal();
   else if( op == "ew_abs" ) Y->ElementWiseAbs();
   else if( op == "ew_sqrt" ) Y->ElementWiseSqrt();
   else if( op == "ew_sgn" ) Y->ElementWiseSgn();
   else if( op == "add_two_vectors" ) Y->AddTwoVectors(a, *X1, b, *X2, c);
   else if( op == "frac_to_bound" ) *out = Y->FracToBound(*X1, a);   // this = x, X1 = delta, a = tau
   else if( op == "add_vector_quotient" ) Y->AddVectorQuotient(a, *X1, *X2, c);   // X1 = z, X2 = s
   else return 1;
   read_back(*Y, y, hy, sy);
   return 0;
}

// ExpansionMatrix P (nrows x ncols, exp_pos 0-based).  which: 0 MultVector (x: ncols -> y: nrows), 1 TransMultVector,
// 2 AddMSinvZ (S = x1, Z = x2: ncols; X = y: nrows), 3 SinvBlrmZMTdBr (S = x1, R = x2, Z = x3: ncols, D = x4: nrows; X = y: ncols)
int vecref_expansion(int which, int nrows, int ncols, const int* exp_pos, double alpha, double beta,
                     const double* x1, int h1, double s1, const double* x2, int h2, double s2,
                     const double* x3, int h3, double s3, const double* x4, int h4, double s4,
                     double* y, int* hy, double* sy)
{
   SmartPtr<ExpansionMatrixSpace> ms = new ExpansionMatrixSpace(nrows, ncols, exp_pos, 0);
   SmartPtr<ExpansionMatrix> P = ms->MakeNewExpansionMatrix();
   SmartPtr<DenseVectorSpace> big = new DenseVectorSpace(nrows);
   SmartPtr<DenseVectorSpace> small = new DenseVectorSpace(ncols);
   if( which == 0 )
   {
      SmartPtr<DenseVector> X = make(*small, x1, h1, s1), Y = make(*big, y, *hy, *sy);
      P->MultVector(alpha, *X, beta, *Y);
      read_back(*Y, y, hy, sy);
   }
   else if( which == 1 )
   {
      SmartPtr<DenseVector> X = make(*big, x1, h1, s1), Y = make(*small, y, *hy, *sy);
      P->TransMultVector(alpha, *X, beta, *Y);
      read_back(*Y, y, hy, sy);
   }
   else if( which == 2 )
   {
      SmartPtr<DenseVector> S = make(*small, x1, h1, s1), Z = make(*small, x2, h2, s2), X = make(*big, y, *hy, *sy);
      P->AddMSinvZ(alpha, *S, *Z, *X);
      read_back(*X, y, hy, sy);
   }
   else if( which == 3 )
   {
      SmartPtr<DenseVector> S = make(*small, x1, h1, s1), R = make(*small, x2, h2, s2), Z = make(*small, x3, h3, s3),
                            D = make(*big, x4, h4, s4), X = make(*small, y, *hy, *sy);
      P->SinvBlrmZMTdBr(alpha, *S, *R, *Z, *D, *X);
      read_back(*X, y, hy, sy);
   }
   else return 1;
   return 0;
}

// triplet matrices (1-based irow / jcol).  symmetric: SymTMatrix (nrows == ncols); trans: TransMultVector
int vecref_tmat(int symmetric, int trans, int nrows, int ncols, int nnz, const int* irow, const int* jcol,
                const double* values, double alpha, double beta, const double* x, int hx, double sx,
                double* y, int* hy, double* sy)
{
   const int nin = trans ? nrows : ncols, nout = trans ? ncols : nrows;
   SmartPtr<DenseVectorSpace> in = new DenseVectorSpace(nin);
   SmartPtr<DenseVectorSpace> outs = new DenseVectorSpace(nout);
   SmartPtr<DenseVector> X = make(*in, x, hx, sx), Y = make(*outs, y, *hy, *sy);
   if( symmetric )
   {
      SmartPtr<SymTMatrixSpace> ms = new SymTMatrixSpace(nrows, nnz, irow, jcol);
      SmartPtr<SymTMatrix> A = ms->MakeNewSymTMatrix();
      if( nnz > 0 ) A->SetValues(values);
      if( trans ) A->TransMultVector(alpha, *X, beta, *Y); else A->MultVector(alpha, *X, beta, *Y);
   }
   else
   {
      SmartPtr<GenTMatrixSpace> ms = new GenTMatrixSpace(nrows, ncols, nnz, irow, jcol);
      SmartPtr<GenTMatrix> A = ms->MakeNewGenTMatrix();
      if( nnz > 0 ) A->SetValues(values);
      if( trans ) A->TransMultVector(alpha, *X, beta, *Y); else A->MultVector(alpha, *X, beta, *Y);
   }
   read_back(*Y, y, hy, sy);
   return 0;
}


// The augmented system exactly as StdAugSystemSolver composes it (reference src/Algorithm/IpStdAugSystemSolver.cpp:232-430)
// from W (SymTMatrix), J_c, J_d (GenTMatrix) and the diagonal vectors, flattened by TripletHelper::FillRowCol / FillValues
// (src/LinAlg/TMatrices/IpTripletHelper.cpp:805-873) -- what TSymLinearSolver hands to the linear solver.
// D_* may be NULL (then only delta).  irn/jcn/vals: capacity nnz_w + n_x + n_s + nnz_jc + n_c + nnz_jd + n_s + n_s.
int vecref_augsys_fill(int n_x, int n_s, int n_c,
                       int nnz_w, const int* w_i, const int* w_j, const double* w_v, double W_factor,
                       int nnz_jc, const int* jc_i, const int* jc_j, const double* jc_v,
                       int nnz_jd, const int* jd_i, const int* jd_j, const double* jd_v,
                       const double* D_x, double delta_x, const double* D_s, double delta_s,
                       const double* D_c, double delta_c, const double* D_d, double delta_d,
                       int* irn, int* jcn, double* vals)
{
   const int n_d = n_s;
   SmartPtr<SymTMatrixSpace> wsp = new SymTMatrixSpace(n_x, nnz_w, w_i, w_j);
   SmartPtr<SymTMatrix> W = wsp->MakeNewSymTMatrix();
   if( nnz_w > 0 ) W->SetValues(w_v);
   SmartPtr<GenTMatrixSpace> jcsp = new GenTMatrixSpace(n_c, n_x, nnz_jc, jc_i, jc_j);
   SmartPtr<GenTMatrix> Jc = jcsp->MakeNewGenTMatrix();
   if( nnz_jc > 0 ) Jc->SetValues(jc_v);
   SmartPtr<GenTMatrixSpace> jdsp = new GenTMatrixSpace(n_d, n_x, nnz_jd, jd_i, jd_j);
   SmartPtr<GenTMatrix> Jd = jdsp->MakeNewGenTMatrix();
   if( nnz_jd > 0 ) Jd->SetValues(jd_v);
   SmartPtr<DenseVectorSpace> vx = new DenseVectorSpace(n_x), vs = new DenseVectorSpace(n_s), vc = new DenseVectorSpace(n_c),
                              vd = new DenseVectorSpace(n_d);
   const int total = n_x + n_s + n_c + n_d;
   SmartPtr<CompoundSymMatrixSpace> asp = new CompoundSymMatrixSpace(4, total);
   asp->SetBlockDim(0, n_x); asp->SetBlockDim(1, n_s); asp->SetBlockDim(2, n_c); asp->SetBlockDim(3, n_d);
   SmartPtr<DiagMatrixSpace> dx = new DiagMatrixSpace(n_x), ds = new DiagMatrixSpace(n_s), dc = new DiagMatrixSpace(n_c),
                             dd = new DiagMatrixSpace(n_d);
   SmartPtr<SumSymMatrixSpace> sx = new SumSymMatrixSpace(n_x, 2);
   sx->SetTermSpace(0, *wsp); sx->SetTermSpace(1, *dx);
   SmartPtr<IdentityMatrixSpace> isp = new IdentityMatrixSpace(n_s);
   asp->SetCompSpace(0, 0, *sx); asp->SetCompSpace(1, 1, *ds); asp->SetCompSpace(2, 0, *jcsp); asp->SetCompSpace(2, 2, *dc);
   asp->SetCompSpace(3, 0, *jdsp); asp->SetCompSpace(3, 1, *isp); asp->SetCompSpace(3, 3, *dd);
   SmartPtr<CompoundSymMatrix> A = asp->MakeNewCompoundSymMatrix();
   struct Mk {
      static SmartPtr<Vector> diag(const DenseVectorSpace& sp, const double* D, double delta)
      {  // the D + delta logic of CreateAugmentedSystem (:344-366)
         SmartPtr<DenseVector> t = sp.MakeNewDenseVector();
         if( D ) { if( sp.Dim() > 0 ) t->SetValues(D); else t->Set(0.); if( delta != 0. ) t->AddScalar(delta); }
         else t->Set(delta);
         return GetRawPtr(t);
      }
   };
   SmartPtr<SumSymMatrix> sumx = sx->MakeNewSumSymMatrix();
   sumx->SetTerm(0, W_factor, *W);
   SmartPtr<DiagMatrix> mdx = dx->MakeNewDiagMatrix(); mdx->SetDiag(*Mk::diag(*vx, D_x, delta_x));
   sumx->SetTerm(1, 1.0, *mdx);
   A->SetComp(0, 0, *sumx);
   SmartPtr<DiagMatrix> mds = ds->MakeNewDiagMatrix(); mds->SetDiag(*Mk::diag(*vs, D_s, delta_s)); A->SetComp(1, 1, *mds);
   A->SetComp(2, 0, *Jc);
   SmartPtr<DiagMatrix> mdc = dc->MakeNewDiagMatrix(); mdc->SetDiag(*Mk::diag(*vc, D_c, -delta_c)); A->SetComp(2, 2, *mdc);
   A->SetComp(3, 0, *Jd);
   SmartPtr<IdentityMatrix> id = isp->MakeNewIdentityMatrix(); id->SetFactor(-1.0); A->SetComp(3, 1, *id);
   SmartPtr<DiagMatrix> mdd = dd->MakeNewDiagMatrix(); mdd->SetDiag(*Mk::diag(*vd, D_d, -delta_d)); A->SetComp(3, 3, *mdd);
   const int nnz = TripletHelper::GetNumberEntries(*A);
   TripletHelper::FillRowCol(nnz, *A, irn, jcn);
   TripletHelper::FillValues(nnz, *A, vals);
   return nnz;
}

}
