// Ipopt-side binding of the b200ldlt C ABI: a SparseSymLinearSolverInterface whose eight virtuals
// forward 1:1 to the extern "C" entry points of include/b200ldlt.h.
//
// Mirrors what the reference's vendor adapters do (e.g. MumpsSolverInterface,
// reference src/Algorithm/LinearSolvers/IpMumpsSolverInterface.{hpp,cpp}); registration needs no change
// to Ipopt: new AlgorithmBuilder(new StdAugSystemSolver(*new TSymLinearSolver(iface, NULL)), "b200-ldlt")
// with linear_solver=custom (IpAlgBuilder.hpp:55-58, IpAlgBuilder.cpp:576-584).
#ifndef B200_LDLT_SOLVER_INTERFACE_HPP
#define B200_LDLT_SOLVER_INTERFACE_HPP

#include <string>
#include <vector>

#include "IpSparseSymLinearSolverInterface.hpp"

namespace Ipopt
{

/** Function table of a backend exposing the b200ldlt-shaped C ABI (the GPU product, or the CPU oracle in tests). */
struct LdltBackend
{
   const char* name;
   void* (*create)(double pivtol, double pivtolmax, int scaling, int verbose, int leaf_k);
   void (*destroy)(void*);
   int (*analyse)(void*, int, int, const int*, const int*);
   double* (*values_ptr)(void*);
   int (*factor)(void*, int, int, int*);
   int (*solve)(void*, int, double*);
   int (*num_neg)(void*);
   int (*increase_quality)(void*);
   /** re-factor the matrix kept by the backend (NULL: use the CALL_AGAIN protocol instead) */
   int (*refactor)(void*, int, int, int*);
   int (*set_pivtol)(void*, double, double);   // optional (NULL: not supported)
};

/** the B200 product backend (libb200ldlt.so) */
const LdltBackend* GetB200LdltBackend();

class B200LdltSolverInterface: public SparseSymLinearSolverInterface
{
public:
   B200LdltSolverInterface(const LdltBackend* backend);
   virtual ~B200LdltSolverInterface();

   virtual bool InitializeImpl(const OptionsList& options, const std::string& prefix);
   virtual ESymSolverStatus InitializeStructure(Index dim, Index nonzeros, const Index* ia, const Index* ja);
   virtual Number* GetValuesArrayPtr();
   virtual ESymSolverStatus MultiSolve(bool new_matrix, const Index* ia, const Index* ja, Index nrhs,
                                       Number* rhs_vals, bool check_NegEVals, Index numberOfNegEVals);
   virtual Index NumberOfNegEVals() const;
   virtual bool IncreaseQuality();
   virtual bool ProvidesInertia() const { return true; }
   virtual EMatrixFormat MatrixFormat() const { return Triplet_Format; }

   static void RegisterOptions(SmartPtr<RegisteredOptions> roptions);

   /** call statistics (host wall clock around the C-ABI calls, i.e. including H2D/D2H) */
   struct Stats
   {
      int n_factor, n_solve, n_rhs, n_singular, n_wrong_inertia;
      int n_analyse;   // InitializeStructure calls that handed a structure to the backend (not the kept ones of a warm start)
      double t_factor, t_solve, t_first_factor;
      int dim, nonzeros;
   };
   const Stats& GetStats() const { return stats_; }
   /** dump (pattern, values, rhs, solution) of the k-th factorisation to <prefix>_<k>.bin (k in dump list) */
   void SetDump(const std::string& prefix, const std::vector<int>& which) { dump_prefix_ = prefix; dump_which_ = which; }

private:
   const LdltBackend* be_;
   void* h_;
   Index dim_, nonzeros_;
   const Index* ia_;
   const Index* ja_;
   Index negevals_;
   bool initialized_, pivtol_changed_, have_factors_;
   Number pivtol_, pivtolmax_;
   Index scaling_, verbose_, leaf_k_;
   bool warm_start_same_structure_;
   Stats stats_;
   std::string dump_prefix_;
   std::vector<int> dump_which_;
   std::vector<Number> dump_vals_;
   bool dump_pending_;
   void DumpSystem(int k, const Number* rhs, const Number* sol, Index nrhs);
};

} // namespace Ipopt
#endif
