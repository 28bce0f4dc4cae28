// See B200LdltSolverInterface.hpp.  Behaviour (status mapping, CALL_AGAIN / IncreaseQuality state machine,
// timing-statistics hooks) follows the reference adapters it replaces:
//   IpMumpsSolverInterface.cpp:191-245 (InitializeImpl), :247-306 (MultiSolve), :349-383 (InitializeStructure),
//   :592-610 (IncreaseQuality); timing hooks as IpSpralSolverInterface.cpp:573-591,676-686.
#include "B200LdltSolverInterface.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>

#include "../../include/b200ldlt.h"
#include "IpIpoptData.hpp"
#include "IpTimingStatistics.hpp"

namespace Ipopt
{

static double wall_now()
{
   using namespace std::chrono;
   return duration<double>(steady_clock::now().time_since_epoch()).count();
}

// ---- product backend glue -------------------------------------------------------------------------------
static void* b200_create(double pivtol, double pivtolmax, int scaling, int verbose, int leaf_k)
{
   b200ldlt_options o;
   b200ldlt_default_options(&o);
   if( pivtol > 0 ) o.pivtol = pivtol;
   if( pivtolmax > 0 ) o.pivtolmax = pivtolmax;
   if( scaling >= 0 ) o.scaling = scaling;
   if( leaf_k > 0 ) o.leaf_k = leaf_k;
   o.verbose = verbose;
   return b200ldlt_create(&o);
}
static void b200_destroy(void* h) { b200ldlt_destroy((b200ldlt_handle) h); }
static int b200_analyse(void* h, int d, int nz, const int* ia, const int* ja) { return b200ldlt_analyse((b200ldlt_handle) h, d, nz, ia, ja); }
static double* b200_values(void* h) { return b200ldlt_values_ptr((b200ldlt_handle) h); }
static int b200_factor(void* h, int c, int e, int* n) { return b200ldlt_factor((b200ldlt_handle) h, c, e, n); }
static int b200_solve(void* h, int nrhs, double* r) { return b200ldlt_solve((b200ldlt_handle) h, nrhs, r); }
static int b200_numneg(void* h) { return b200ldlt_num_neg((b200ldlt_handle) h); }
static int b200_incq(void* h) { return b200ldlt_increase_quality((b200ldlt_handle) h); }
static int b200_refactor(void* h, int c, int e, int* n) { return b200ldlt_refactor((b200ldlt_handle) h, c, e, n); }
static int b200_set_pivtol(void* h, double u, double umax) { return b200ldlt_set_pivtol((b200ldlt_handle) h, u, umax); }

const LdltBackend* GetB200LdltBackend()
{
   static const LdltBackend be = {"b200-ldlt", b200_create, b200_destroy, b200_analyse, b200_values, b200_factor,
                                  b200_solve, b200_numneg, b200_incq, b200_refactor, b200_set_pivtol};
   return &be;
}

// ---------------------------------------------------------------------------------------------------------
B200LdltSolverInterface::B200LdltSolverInterface(const LdltBackend* backend)
   : be_(backend), h_(NULL), dim_(0), nonzeros_(0), ia_(NULL), ja_(NULL), negevals_(-1), initialized_(false),
     pivtol_changed_(false), have_factors_(false), pivtol_(1e-8), pivtolmax_(1e-4), scaling_(2), verbose_(0),
     leaf_k_(0), warm_start_same_structure_(false), dump_pending_(false)
{
   memset(&stats_, 0, sizeof(stats_));
}

B200LdltSolverInterface::~B200LdltSolverInterface()
{
   if( h_ ) be_->destroy(h_);
}

void B200LdltSolverInterface::RegisterOptions(SmartPtr<RegisteredOptions> roptions)
{
   roptions->AddBoundedNumberOption("b200_pivtol", "Pivot threshold u of the B200 LDL^T backend.", 0.0, true, 1.0, true,
                                    1e-8, "Same meaning as ma57_pivtol.");
   roptions->AddBoundedNumberOption("b200_pivtolmax", "Maximum pivot threshold (IncreaseQuality cap).", 0.0, true, 1.0,
                                    true, 1e-4, "Same meaning as ma57_pivtolmax.");
   roptions->AddLowerBoundedIntegerOption("b200_scaling", "Symmetric equilibration sweeps done inside the backend.", 0, 2,
                                          "0 disables the internal scaling.");
   roptions->AddLowerBoundedIntegerOption("b200_verbose", "Verbosity of the backend (stderr).", 0, 0, "");
   roptions->AddLowerBoundedIntegerOption("b200_leaf_k", "Elimination subtrees up to this many columns become one front.", 0,
                                          0, "0 = backend default.");
}

bool B200LdltSolverInterface::InitializeImpl(const OptionsList& options, const std::string& prefix)
{
   options.GetNumericValue("b200_pivtol", pivtol_, prefix);
   if( options.GetNumericValue("b200_pivtolmax", pivtolmax_, prefix) )
   {
      ASSERT_EXCEPTION(pivtolmax_ >= pivtol_, OPTION_INVALID, "Option \"b200_pivtolmax\": This value must be between b200_pivtol and 1.");
   }
   else
   {
      pivtolmax_ = Max(pivtolmax_, pivtol_);
   }
   options.GetIntegerValue("b200_scaling", scaling_, prefix);
   options.GetIntegerValue("b200_verbose", verbose_, prefix);
   options.GetIntegerValue("b200_leaf_k", leaf_k_, prefix);
   options.GetBoolValue("warm_start_same_structure", warm_start_same_structure_, prefix);

   // reset the per-(re)optimisation state, keep the symbolic data if the user promises the same structure
   // (cf. IpMumpsSolverInterface.cpp:227-236)
   pivtol_changed_ = false;
   if( !warm_start_same_structure_ || !h_ )
   {
      if( h_ ) be_->destroy(h_);
      h_ = be_->create(pivtol_, pivtolmax_, scaling_, verbose_, leaf_k_);
      initialized_ = false;
      have_factors_ = false;
      if( !h_ )
      {
         Jnlst().Printf(J_ERROR, J_LINEAR_ALGEBRA, "%s: backend could not be created (no usable CUDA device?)\n", be_->name);
         return false;
      }
   }
   else
   {
      ASSERT_EXCEPTION(initialized_, INVALID_WARMSTART, "B200LdltSolverInterface called with warm_start_same_structure, but the problem is solved for the first time.");
      // kept handle: the freshly read thresholds replace whatever an earlier IncreaseQuality left behind
      // (the reference adapters reset pivtol in every InitializeImpl, IpMumpsSolverInterface.cpp:191-245)
      if( be_->set_pivtol ) be_->set_pivtol(h_, pivtol_, pivtolmax_);
   }
   return true;
}

ESymSolverStatus B200LdltSolverInterface::InitializeStructure(Index dim, Index nonzeros, const Index* ia, const Index* ja)
{
   ESymSolverStatus retval = SYMSOLVER_SUCCESS;
   if( !warm_start_same_structure_ || !initialized_ )
   {
      dim_ = dim;
      nonzeros_ = nonzeros;
      ia_ = ia;
      ja_ = ja;
      int rc = be_->analyse(h_, dim, nonzeros, ia, ja);
      if( rc != 0 )
      {
         return SYMSOLVER_FATAL_ERROR;
      }
      have_factors_ = false;
      ++stats_.n_analyse;
      stats_.dim = dim;
      stats_.nonzeros = nonzeros;
   }
   else
   {
      ASSERT_EXCEPTION(dim_ == dim && nonzeros_ == nonzeros, INVALID_WARMSTART, "B200LdltSolverInterface called with warm_start_same_structure, but the problem size has changed.");
   }
   initialized_ = true;
   return retval;
}

Number* B200LdltSolverInterface::GetValuesArrayPtr()
{
   return be_->values_ptr(h_);
}

ESymSolverStatus B200LdltSolverInterface::MultiSolve(bool new_matrix, const Index* ia, const Index* ja, Index nrhs,
      Number* rhs_vals, bool check_NegEVals, Index numberOfNegEVals)
{
   DBG_ASSERT(initialized_);
   (void) ia; (void) ja;
   // pivtol changed (IncreaseQuality) and Ipopt asks to re-use the factors: re-factor the matrix kept on the
   // device instead of bouncing CALL_AGAIN through TSymLinearSolver (IpMumpsSolverInterface.cpp:265-278).
   bool refactor = false;
   if( pivtol_changed_ )
   {
      pivtol_changed_ = false;
      if( !new_matrix )
      {
         if( be_->refactor == NULL )
         {
            return SYMSOLVER_CALL_AGAIN;
         }
         refactor = true;
      }
   }
   if( new_matrix || refactor )
   {
      if( !dump_which_.empty() )
      {
         dump_pending_ = false;
         for( size_t q = 0; q < dump_which_.size(); ++q )
            if( dump_which_[q] == stats_.n_factor ) dump_pending_ = true;
         if( dump_pending_ && new_matrix ) dump_vals_.assign(be_->values_ptr(h_), be_->values_ptr(h_) + nonzeros_);
      }
      if( HaveIpData() )
      {
         if( stats_.n_factor == 0 ) IpData().TimingStats().LinearSystemSymbolicFactorization().Start();
         else IpData().TimingStats().LinearSystemFactorization().Start();
      }
      double t0 = wall_now();
      int neg = -1;
      int st = refactor ? be_->refactor(h_, check_NegEVals ? 1 : 0, numberOfNegEVals, &neg)
               : be_->factor(h_, check_NegEVals ? 1 : 0, numberOfNegEVals, &neg);
      double dt = wall_now() - t0;
      if( HaveIpData() )
      {
         if( stats_.n_factor == 0 ) IpData().TimingStats().LinearSystemSymbolicFactorization().End();
         else IpData().TimingStats().LinearSystemFactorization().End();
      }
      if( stats_.n_factor == 0 ) stats_.t_first_factor = dt;   // includes the one-off ordering + symbolic phase
      else stats_.t_factor += dt;
      stats_.n_factor++;
      negevals_ = neg;
      have_factors_ = (st == B200LDLT_SUCCESS);
      if( st == B200LDLT_SINGULAR ) { stats_.n_singular++; return SYMSOLVER_SINGULAR; }
      if( st == B200LDLT_WRONG_INERTIA ) { stats_.n_wrong_inertia++; return SYMSOLVER_WRONG_INERTIA; }
      if( st != B200LDLT_SUCCESS )
      {
         Jnlst().Printf(J_ERROR, J_LINEAR_ALGEBRA, "%s: factorisation failed with status %d\n", be_->name, st);
         return SYMSOLVER_FATAL_ERROR;
      }
   }
   if( !have_factors_ ) return SYMSOLVER_FATAL_ERROR;
   std::vector<Number> rhs_copy;
   if( dump_pending_ ) rhs_copy.assign(rhs_vals, rhs_vals + (size_t) dim_ * nrhs);
   if( HaveIpData() ) IpData().TimingStats().LinearSystemBackSolve().Start();
   double t0 = wall_now();
   int st = be_->solve(h_, nrhs, rhs_vals);
   stats_.t_solve += wall_now() - t0;
   stats_.n_solve++;
   stats_.n_rhs += nrhs;
   if( HaveIpData() ) IpData().TimingStats().LinearSystemBackSolve().End();
   if( dump_pending_ )
   {
      DumpSystem(stats_.n_factor - 1, &rhs_copy[0], rhs_vals, nrhs);
      dump_pending_ = false;
   }
   return st == B200LDLT_SUCCESS ? SYMSOLVER_SUCCESS : SYMSOLVER_FATAL_ERROR;
}

Index B200LdltSolverInterface::NumberOfNegEVals() const
{
   return negevals_;
}

bool B200LdltSolverInterface::IncreaseQuality()
{
   if( !be_->increase_quality(h_) ) return false;
   pivtol_changed_ = true;
   Jnlst().Printf(J_DETAILED, J_LINEAR_ALGEBRA, "Increasing pivot tolerance of %s.\n", be_->name);
   return true;
}

// binary dump: int32 dim, nnz, nrhs, neg; int32 irn[nnz], jcn[nnz]; f64 vals[nnz], rhs[dim*nrhs], sol[dim*nrhs]
void B200LdltSolverInterface::DumpSystem(int k, const Number* rhs, const Number* sol, Index nrhs)
{
   char name[1024];
   snprintf(name, sizeof(name), "%s_%d.bin", dump_prefix_.c_str(), k);
   FILE* fp = fopen(name, "wb");
   if( !fp ) return;
   int hdr[4] = {dim_, nonzeros_, nrhs, negevals_};
   fwrite(hdr, sizeof(int), 4, fp);
   fwrite(ia_, sizeof(int), nonzeros_, fp);
   fwrite(ja_, sizeof(int), nonzeros_, fp);
   fwrite(&dump_vals_[0], sizeof(double), nonzeros_, fp);
   fwrite(rhs, sizeof(double), (size_t) dim_ * nrhs, fp);
   fwrite(sol, sizeof(double), (size_t) dim_ * nrhs, fp);
   fclose(fp);
}

} // namespace Ipopt
