// dp45_host.h -- host-side Dormand-Prince 5(4) driver used only by the stand-in
// host's recombination solve (the reference calls its RungeKuttaDP45 from
// recfast.f90:758-785).  Control flow follows fortran/RungeKuttaDP45.f90:94-408
// for the subset CAMB uses: error_control = mixed, no interrupts, optional FSAL.
// The GPU path has its own warp-level implementation in csrc/cuda/evolve.cu.
#pragma once
#include <cmath>
#include <functional>
#include <limits>
#include <vector>

namespace cbh {

struct DP45Host {
    // persistent state (RungeKuttaDP45Settings, RungeKuttaDP45.f90:14-73)
    bool reuse_first_derivative = false;
    double max_step_size = 0;
    double current_step_size = 0, estimated_error = 0, previous_xend = 0;
    bool xend_reached = false, first_derivative_valid = false;
    int successive_failures = 0;
    long function_evaluations = 0, successful_steps = 0;
    int n = 0;
    std::vector<double> w;  // n x 9, column-major

    using rhs = std::function<void(double, const double*, double*)>;

    // returns ind (3 ok, negative on failure)
    int integrate(const rhs& fcn, int n_, double& x, double* y, double xend, double tol, int ind) {
        n = n_;
        if ((int)w.size() < 9 * n) w.assign(9 * n, 0.0);
        auto W = [&](int col) { return w.data() + (size_t)(col - 1) * n; };
        const double eps = std::numeric_limits<double>::epsilon();
        const double tiny = std::numeric_limits<double>::min();
        if (ind == 1 || ind == 2) {
            if (ind == 1) {
                bool keep = reuse_first_derivative;
                *this = DP45Host();
                reuse_first_derivative = keep;
                n = n_;
                w.assign(9 * n, 0.0);
            }
            previous_xend = x;
            xend_reached = false;
            first_derivative_valid = false;
            successful_steps = 0; successive_failures = 0; function_evaluations = 0;
        } else if (ind == 3) {
            if (xend_reached && (x != previous_xend || xend == previous_xend)) return -9;
            xend_reached = false;
        } else return -9;

        for (;;) {
            if (ind != 6) {
                if (reuse_first_derivative && first_derivative_valid) first_derivative_valid = false;
                else { fcn(x, y, W(1)); function_evaluations++; }
            }
            double temp = 0;
            for (int k = 0; k < n; k++) temp = std::fmax(temp, std::fabs(y[k]));
            double wny = std::fmin(temp, 1.0);
            double hmin = 10.0 * std::fmax(tiny, eps * std::fmax(wny / tol, std::fabs(x)));
            double hmax = max_step_size != 0 ? std::fabs(max_step_size) : 20.0;
            if (hmin > hmax) return -2;
            if (ind <= 2) current_step_size = hmax * std::pow(tol, 0.2);
            else if (successive_failures <= 1) {
                temp = 2.0 * current_step_size;
                if (tol < std::pow(2.0 / 0.9, 5) * estimated_error)
                    temp = 0.9 * std::pow(tol / estimated_error, 0.2) * current_step_size;
                current_step_size = std::fmax(temp, 0.5 * current_step_size);
            } else current_step_size = 0.5 * current_step_size;
            current_step_size = std::fmin(current_step_size, hmax);
            current_step_size = std::fmax(current_step_size, hmin);

            temp = std::fabs(xend - x);
            double trial_x;
            if (current_step_size < temp && (temp > 1.2 * current_step_size || temp > hmax)) {
                current_step_size = std::fmin(current_step_size, 0.5 * temp);
                trial_x = x + std::copysign(current_step_size, xend - x);
            } else { current_step_size = temp; trial_x = xend; }
            double h = trial_x - x;
            double *k1 = W(1), *k2 = W(2), *k3 = W(3), *k4 = W(4), *k5 = W(5), *k6 = W(6), *k7 = W(7), *yt = W(9);
            for (int k = 0; k < n; k++) yt[k] = y[k] + h * (1.0 / 5.0) * k1[k];
            fcn(x + h / 5.0, yt, k2);
            for (int k = 0; k < n; k++) yt[k] = y[k] + h * ((3.0 / 40.0) * k1[k] + (9.0 / 40.0) * k2[k]);
            fcn(x + h * (3.0 / 10.0), yt, k3);
            for (int k = 0; k < n; k++)
                yt[k] = y[k] + h * ((44.0 / 45.0) * k1[k] + (-56.0 / 15.0) * k2[k] + (32.0 / 9.0) * k3[k]);
            fcn(x + h * (4.0 / 5.0), yt, k4);
            for (int k = 0; k < n; k++)
                yt[k] = y[k] + h * ((19372.0 / 6561.0) * k1[k] + (-25360.0 / 2187.0) * k2[k] +
                                    (64448.0 / 6561.0) * k3[k] + (-212.0 / 729.0) * k4[k]);
            fcn(x + h * (8.0 / 9.0), yt, k5);
            for (int k = 0; k < n; k++)
                yt[k] = y[k] + h * ((9017.0 / 3168.0) * k1[k] + (-355.0 / 33.0) * k2[k] + (46732.0 / 5247.0) * k3[k] +
                                    (49.0 / 176.0) * k4[k] + (-5103.0 / 18656.0) * k5[k]);
            fcn(x + h, yt, k6);
            for (int k = 0; k < n; k++)
                yt[k] = y[k] + h * ((35.0 / 384.0) * k1[k] + (500.0 / 1113.0) * k3[k] + (125.0 / 192.0) * k4[k] +
                                    (-2187.0 / 6784.0) * k5[k] + (11.0 / 84.0) * k6[k]);
            fcn(x + h, yt, k7);
            function_evaluations += 6;
            temp = 0;
            for (int k = 0; k < n; k++) {
                double e = (71.0 / 57600.0) * k1[k] + (-71.0 / 16695.0) * k3[k] + (71.0 / 1920.0) * k4[k] +
                           (-17253.0 / 339200.0) * k5[k] + (22.0 / 525.0) * k6[k] + (-1.0 / 40.0) * k7[k];
                temp = std::fmax(temp, std::fabs(e) / std::fmax(1.0, std::fabs(y[k])));
            }
            estimated_error = temp * current_step_size;
            ind = estimated_error > tol ? 6 : 5;
            if (ind == 5) {
                x = trial_x;
                for (int k = 0; k < n; k++) y[k] = yt[k];
                successful_steps++;
                successive_failures = 0;
                if (x == xend) {
                    previous_xend = xend;
                    xend_reached = true;
                    if (reuse_first_derivative) {
                        for (int k = 0; k < n; k++) k1[k] = k7[k];
                        first_derivative_valid = true;
                    }
                    return 3;
                }
                for (int k = 0; k < n; k++) k1[k] = k7[k];
                ind = 6;
            } else {
                successive_failures++;
                if (current_step_size <= hmin) return -3;
            }
        }
    }
};

}  // namespace cbh
