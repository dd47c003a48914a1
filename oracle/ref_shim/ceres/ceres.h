// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Stand-in for the slice of Ceres Solver 2.0 (README.md:21 of the reference; third-party, not under /root/reference)
// that pose_estimation.cpp:14-47,100-143 uses: Jet-based AutoDiffCostFunction<F, 2, 3, 3>, Problem::AddResidualBlock
// with a NULL loss, Solver::Options {linear_solver_type, gradient/function/parameter_tolerance}, Solve().  Solve()
// restates Ceres' published default minimiser for such a problem: trust-region Levenberg-Marquardt
// (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc) with Jacobi scaling, initial radius 1e4, LM diagonal
// clamped to [1e-6, 1e32], radius update r / max(1/3, 1 - (2 rho - 1)^3) on success and r / 2, 4, 8.. on failure,
// min_relative_decrease 1e-3, at most 50 iterations.  The 6x6 damped normal equations are solved by Cholesky (what
// DENSE_SCHUR reduces to for two dense parameter blocks).  The minimiser's end point, not its path, is what the pose
// bar (1e-4 rad / 1e-4 |t|) looks at.
#pragma once
#ifndef CTAG_REF_SHIM_CERES_H
#define CTAG_REF_SHIM_CERES_H

#include <cmath>
#include <cstddef>
#include <memory>
#include <string>
#include <vector>

namespace ceres {

template <typename T, int N> struct Jet {
    T a;
    T v[N];
    Jet() : a(0) { for (int i = 0; i < N; i++) v[i] = 0; }
    Jet(const T& s) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; }  // NOLINT: implicit like ceres::Jet
    Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = 0; v[k] = 1; }
    Jet& operator+=(const Jet& y) { a += y.a; for (int i = 0; i < N; i++) v[i] += y.v[i]; return *this; }
    Jet& operator-=(const Jet& y) { a -= y.a; for (int i = 0; i < N; i++) v[i] -= y.v[i]; return *this; }
    Jet& operator*=(const Jet& y) { *this = *this * y; return *this; }
    Jet& operator/=(const Jet& y) { *this = *this / y; return *this; }
};
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f) { return f; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) {
    Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h;
}
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] + g.v[i]; return h;
}
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; i++) h.v[i] = f.v[i] - g.v[i]; return h;
}
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, const Jet<T, N>& g) {
    Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; i++) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h;
}
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g) {
    const T gi = T(1.0) / g.a, q = f.a * gi;
    Jet<T, N> h; h.a = q; for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h;
}
#define CTAG_SHIM_JET_SCALAR(op)                                                                                       \
    template <typename T, int N> inline Jet<T, N> operator op(const Jet<T, N>& f, T s) { return f op Jet<T, N>(s); }   \
    template <typename T, int N> inline Jet<T, N> operator op(T s, const Jet<T, N>& f) { return Jet<T, N>(s) op f; }
CTAG_SHIM_JET_SCALAR(+)
CTAG_SHIM_JET_SCALAR(-)
CTAG_SHIM_JET_SCALAR(*)
CTAG_SHIM_JET_SCALAR(/)
#undef CTAG_SHIM_JET_SCALAR
#define CTAG_SHIM_JET_CMP(op)                                                                                          \
    template <typename T, int N> inline bool operator op(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a op g.a; } \
    template <typename T, int N> inline bool operator op(const Jet<T, N>& f, T s) { return f.a op s; }                 \
    template <typename T, int N> inline bool operator op(T s, const Jet<T, N>& f) { return s op f.a; }
CTAG_SHIM_JET_CMP(<)
CTAG_SHIM_JET_CMP(<=)
CTAG_SHIM_JET_CMP(>)
CTAG_SHIM_JET_CMP(>=)
CTAG_SHIM_JET_CMP(==)
CTAG_SHIM_JET_CMP(!=)
#undef CTAG_SHIM_JET_CMP
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) {
    Jet<T, N> h; h.a = std::sqrt(f.a); const T d = T(1.0) / (T(2.0) * h.a);
    for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h;
}
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) {
    Jet<T, N> h; h.a = std::sin(f.a); const T d = std::cos(f.a);
    for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h;
}
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) {
    Jet<T, N> h; h.a = std::cos(f.a); const T d = -std::sin(f.a);
    for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h;
}
class LossFunction;

class CostFunction {
public:
    virtual ~CostFunction() {}
    // jacobians[b] (may be NULL) is num_residuals x block_size(b), row-major
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
    int num_residuals() const { return num_residuals_; }
    const std::vector<int>& parameter_block_sizes() const { return sizes_; }
protected:
    int num_residuals_ = 0;
    std::vector<int> sizes_;
};

template <typename Functor, int kNumResiduals, int N0, int N1> class AutoDiffCostFunction : public CostFunction {
public:
    explicit AutoDiffCostFunction(Functor* f) : functor_(f) {
        num_residuals_ = kNumResiduals;
        sizes_ = {N0, N1};
    }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override {
        if (!jacobians) return (*functor_)(parameters[0], parameters[1], residuals);
        typedef Jet<double, N0 + N1> J;
        J x0[N0], x1[N1], out[kNumResiduals];
        for (int i = 0; i < N0; i++) x0[i] = J(parameters[0][i], i);
        for (int i = 0; i < N1; i++) x1[i] = J(parameters[1][i], N0 + i);
        if (!(*functor_)(x0, x1, out)) return false;
        for (int r = 0; r < kNumResiduals; r++) {
            residuals[r] = out[r].a;
            if (jacobians[0]) for (int i = 0; i < N0; i++) jacobians[0][r * N0 + i] = out[r].v[i];
            if (jacobians[1]) for (int i = 0; i < N1; i++) jacobians[1][r * N1 + i] = out[r].v[N0 + i];
        }
        return true;
    }
private:
    std::unique_ptr<Functor> functor_;
};

class Problem {
public:
    struct Block {
        std::unique_ptr<CostFunction> cost;
        double* x0;
        double* x1;
    };
    Problem() {}
    Problem(const Problem&) = delete;
    void AddResidualBlock(CostFunction* cost, LossFunction* loss, double* x0, double* x1) {
        (void)loss;  // the reference passes NULL (pose_estimation.cpp:139)
        Block b; b.cost.reset(cost); b.x0 = x0; b.x1 = x1;
        blocks_.push_back(std::move(b));
    }
    int NumResidualBlocks() const { return (int)blocks_.size(); }
    const std::vector<Block>& blocks() const { return blocks_; }
private:
    std::vector<Block> blocks_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

class Solver {
public:
    struct Options {
        LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
        int max_num_iterations = 50;
        double gradient_tolerance = 1e-10;
        double function_tolerance = 1e-6;
        double parameter_tolerance = 1e-8;
        double initial_trust_region_radius = 1e4;
        double max_trust_region_radius = 1e16;
        double min_trust_region_radius = 1e-32;
        double min_relative_decrease = 1e-3;
        double min_lm_diagonal = 1e-6;
        double max_lm_diagonal = 1e32;
        int max_num_consecutive_invalid_steps = 5;
        bool jacobi_scaling = true;
        bool minimizer_progress_to_stdout = false;
    };
    struct Summary {
        TerminationType termination_type = FAILURE;
        double initial_cost = -1, final_cost = -1;
        int num_successful_steps = 0, num_unsuccessful_steps = 0;
        std::string message;
        std::string BriefReport() const;
    };
};

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);

}  // namespace ceres

#endif
