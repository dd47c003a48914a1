// TEST INFRASTRUCTURE ONLY -- never linked into the product.
// ceres::Solve for the stand-in in ceres/ceres.h: trust-region Levenberg-Marquardt as published for Ceres 2.0's
// default minimiser (see the header comment).  All parameter blocks of the problem are gathered into one dense vector.
#include "ceres/ceres.h"

#include <algorithm>
#include <cstdio>
#include <map>

namespace ceres {

namespace {

struct Dense {
    std::vector<double*> blocks;   // distinct parameter blocks in first-use order
    std::vector<int> sizes, offs;
    int np = 0, nr = 0;
};

int block_index(Dense& d, double* p, int size) {
    for (size_t i = 0; i < d.blocks.size(); i++) if (d.blocks[i] == p) return (int)i;
    d.blocks.push_back(p); d.sizes.push_back(size); d.offs.push_back(d.np); d.np += size;
    return (int)d.blocks.size() - 1;
}

// residuals r (nr), optional dense row-major Jacobian J (nr x np) at x
bool evaluate(const Problem& pb, Dense& d, const std::vector<double>& x, std::vector<double>& r, std::vector<double>* J) {
    r.assign(d.nr, 0.0);
    if (J) J->assign((size_t)d.nr * d.np, 0.0);
    int row = 0;
    for (const auto& b : pb.blocks()) {
        const std::vector<int>& sz = b.cost->parameter_block_sizes();
        const int i0 = block_index(d, b.x0, sz[0]), i1 = block_index(d, b.x1, sz[1]);
        const double* params[2] = {&x[d.offs[i0]], &x[d.offs[i1]]};
        const int m = b.cost->num_residuals();
        std::vector<double> j0((size_t)m * sz[0]), j1((size_t)m * sz[1]);
        double* jac[2] = {j0.data(), j1.data()};
        if (!b.cost->Evaluate(params, &r[row], J ? jac : nullptr)) return false;
        if (J)
            for (int k = 0; k < m; k++) {
                for (int c = 0; c < sz[0]; c++) (*J)[(size_t)(row + k) * d.np + d.offs[i0] + c] += j0[(size_t)k * sz[0] + c];
                for (int c = 0; c < sz[1]; c++) (*J)[(size_t)(row + k) * d.np + d.offs[i1] + c] += j1[(size_t)k * sz[1] + c];
            }
        row += m;
    }
    return true;
}

// solves (A + diag(dd)) s = b for symmetric positive definite A (n x n) by Cholesky; false if not SPD
bool chol_solve(std::vector<double> A, const std::vector<double>& dd, std::vector<double> b, int n, std::vector<double>& s) {
    for (int i = 0; i < n; i++) A[(size_t)i * n + i] += dd[i];
    for (int j = 0; j < n; j++) {
        double v = A[(size_t)j * n + j];
        for (int k = 0; k < j; k++) v -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
        if (!(v > 0)) return false;
        v = std::sqrt(v);
        A[(size_t)j * n + j] = v;
        for (int i = j + 1; i < n; i++) {
            double t = A[(size_t)i * n + j];
            for (int k = 0; k < j; k++) t -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
            A[(size_t)i * n + j] = t / v;
        }
    }
    for (int i = 0; i < n; i++) { double t = b[i]; for (int k = 0; k < i; k++) t -= A[(size_t)i * n + k] * b[k]; b[i] = t / A[(size_t)i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { double t = b[i]; for (int k = i + 1; k < n; k++) t -= A[(size_t)k * n + i] * b[k]; b[i] = t / A[(size_t)i * n + i]; }
    s = b;
    return true;
}

double sqnorm(const std::vector<double>& v) { double s = 0; for (double a : v) s += a * a; return s; }

}  // namespace

std::string Solver::Summary::BriefReport() const {
    char buf[160];
    std::snprintf(buf, sizeof buf, "shim-ceres: initial cost %.6e, final cost %.6e, %d+%d steps, %s", initial_cost, final_cost,
                  num_successful_steps, num_unsuccessful_steps, message.c_str());
    return buf;
}

void Solve(const Solver::Options& opt, Problem* problem, Solver::Summary* sum) {
    Solver::Summary local;
    Solver::Summary& S = sum ? *sum : local;
    S = Solver::Summary();
    Dense d;
    for (const auto& b : problem->blocks()) {
        const std::vector<int>& sz = b.cost->parameter_block_sizes();
        block_index(d, b.x0, sz[0]); block_index(d, b.x1, sz[1]);
        d.nr += b.cost->num_residuals();
    }
    const int n = d.np;
    if (n == 0 || d.nr == 0) { S.termination_type = CONVERGENCE; S.message = "empty problem"; S.initial_cost = S.final_cost = 0; return; }
    std::vector<double> x(n), r, J, cand(n), rc;
    for (size_t i = 0; i < d.blocks.size(); i++) for (int k = 0; k < d.sizes[i]; k++) x[d.offs[i] + k] = d.blocks[i][k];
    auto store = [&]() { for (size_t i = 0; i < d.blocks.size(); i++) for (int k = 0; k < d.sizes[i]; k++) d.blocks[i][k] = x[d.offs[i] + k]; };
    if (!evaluate(*problem, d, x, r, &J)) { S.message = "initial evaluation failed"; return; }
    double cost = 0.5 * sqnorm(r);
    S.initial_cost = S.final_cost = cost;

    std::vector<double> scale(n, 1.0);
    if (opt.jacobi_scaling)
        for (int c = 0; c < n; c++) { double s = 0; for (int k = 0; k < d.nr; k++) s += J[(size_t)k * n + c] * J[(size_t)k * n + c]; scale[c] = 1.0 / (1.0 + std::sqrt(s)); }
    std::vector<double> A((size_t)n * n), g(n), gs(n), diag(n), lm(n), step(n);
    auto normal_eq = [&]() {  // scaled J: A = Js' Js, gs = Js' r; g = J' r (unscaled gradient)
        for (int a = 0; a < n; a++) {
            double ga = 0;
            for (int k = 0; k < d.nr; k++) ga += J[(size_t)k * n + a] * r[k];
            g[a] = ga; gs[a] = ga * scale[a];
            for (int b = a; b < n; b++) {
                double s = 0;
                for (int k = 0; k < d.nr; k++) s += J[(size_t)k * n + a] * J[(size_t)k * n + b];
                A[(size_t)a * n + b] = A[(size_t)b * n + a] = s * scale[a] * scale[b];
            }
        }
    };
    auto max_norm = [&](const std::vector<double>& v) { double m = 0; for (double a : v) m = std::max(m, std::fabs(a)); return m; };
    normal_eq();
    if (max_norm(g) <= opt.gradient_tolerance) { S.termination_type = CONVERGENCE; S.message = "gradient tolerance"; store(); return; }

    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    int invalid = 0;
    double x_norm = std::sqrt(sqnorm(x));
    S.termination_type = NO_CONVERGENCE; S.message = "max iterations";
    for (int it = 1; it <= opt.max_num_iterations; it++) {
        if (!reuse_diagonal) for (int c = 0; c < n; c++) diag[c] = std::min(std::max(A[(size_t)c * n + c], opt.min_lm_diagonal), opt.max_lm_diagonal);
        for (int c = 0; c < n; c++) lm[c] = diag[c] / radius;  // D^2
        std::vector<double> neg(n);
        for (int c = 0; c < n; c++) neg[c] = -gs[c];
        bool valid = chol_solve(A, lm, neg, n, step);
        double model_change = 0;
        if (valid) {  // -(Js s)'(Js s / 2 + r) = -(s'A s / 2 + s'gs)
            double sAs = 0, sg = 0;
            for (int a = 0; a < n; a++) { sg += step[a] * gs[a]; for (int b = 0; b < n; b++) sAs += step[a] * A[(size_t)a * n + b] * step[b]; }
            model_change = -(0.5 * sAs + sg);
            valid = model_change > 0;
        }
        if (!valid) {
            if (++invalid > opt.max_num_consecutive_invalid_steps) { S.termination_type = FAILURE; S.message = "too many invalid steps"; break; }
            radius /= decrease_factor; decrease_factor *= 2; reuse_diagonal = true;
            S.num_unsuccessful_steps++;
            if (radius < opt.min_trust_region_radius) { S.termination_type = CONVERGENCE; S.message = "min trust region radius"; break; }
            continue;
        }
        invalid = 0;
        double step_sq = 0;
        for (int c = 0; c < n; c++) { double dl = step[c] * scale[c]; cand[c] = x[c] + dl; step_sq += dl * dl; }
        double cand_cost = evaluate(*problem, d, cand, rc, nullptr) ? 0.5 * sqnorm(rc) : std::numeric_limits<double>::max();
        if (std::sqrt(step_sq) <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { S.termination_type = CONVERGENCE; S.message = "parameter tolerance"; break; }
        if (std::fabs(cost - cand_cost) <= opt.function_tolerance * cost) { S.termination_type = CONVERGENCE; S.message = "function tolerance"; break; }
        double rho = (cost - cand_cost) / model_change;
        if (rho > opt.min_relative_decrease) {
            x = cand; x_norm = std::sqrt(sqnorm(x)); cost = cand_cost;
            if (!evaluate(*problem, d, x, r, &J)) { S.termination_type = FAILURE; S.message = "evaluation failed"; break; }
            normal_eq();
            radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
            decrease_factor = 2.0; reuse_diagonal = false;
            S.num_successful_steps++;
            if (max_norm(g) <= opt.gradient_tolerance) { S.termination_type = CONVERGENCE; S.message = "gradient tolerance"; break; }
        } else {
            radius /= decrease_factor; decrease_factor *= 2; reuse_diagonal = true;
            S.num_unsuccessful_steps++;
            if (radius < opt.min_trust_region_radius) { S.termination_type = CONVERGENCE; S.message = "min trust region radius"; break; }
        }
    }
    S.final_cost = cost;
    store();
    if (opt.minimizer_progress_to_stdout) std::printf("%s\n", S.BriefReport().c_str());
}

}  // namespace ceres
